# what the driver runs at round end, in one go: gpu tests, smoke, the default bench line, the reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.txt 2>&1
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
cat gpurun_out/final_pytest_gpu.txt gpurun_out/final_smoke.txt; tail -c 400 gpurun_out/final_bench_ref.json
