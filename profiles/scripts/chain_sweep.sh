# k_att_chain (queue kernel, cw=-1) against k_att_chain_spec with cw warps per chain: batch (1 and 6 plan waves) and one C2 track
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cw in ${W1:--1 1 2}; do echo "waves1 cw=$cw"; python bench.py --no-cpu-baseline --no-e2e --waves 1 --chain-warps $cw 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel']['k_att_chain'])"; done
for cw in ${W6:--1 1 2 4}; do echo "waves6 cw=$cw"; python bench.py --no-cpu-baseline --no-e2e --waves 6 --chain-warps $cw 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel']['k_att_chain'])"; done
for cw in ${C2:--1 1 2 4 8}; do echo "c2 cw=$cw"; AME_CHAIN_WARPS=$cw python profiles/scripts/single_track.py c2 2>&1 | tail -2; done
