# round-1 (e) evidence: gpu tests, default bench, ncu launch list of the bench command, ncu --set full of one 1-wave step
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r01e_pytest_gpu.txt
python bench.py > gpurun_out/r01e_bench_n1.json 2> gpurun_out/r01e_bench_n1.err
python bench.py --no-e2e --no-cpu-baseline --waves 1 > gpurun_out/r01e_bench_1wave.json 2>> gpurun_out/r01e_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r01e_launches_ncu.csv \
    python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/r01e_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_ -s 11 -c 11 -o gpurun_out/r01e_full_1wave \
    python bench.py --no-e2e --no-cpu-baseline --waves 1 --steps 1 --warmup 1 > gpurun_out/r01e_ncu_full.log 2>&1
tail -2 gpurun_out/r01e_ncu_full.log
