# device-resident step and end-to-end step of the bench batch against chain_warps (k_att_chain_spec) and plan waves
dev() { python bench.py --no-cpu-baseline --no-e2e --waves $1 --chain-warps $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('device waves=$1 cw=$2 ms/step', round(d['ms_per_step'],2), 'chain GB/s', d['roofline']['per_kernel']['k_att_chain']['achieved'])"; }
e2e() { python bench.py --no-cpu-baseline --steps 2 --warmup 3 --waves 6 --chain-warps $2 --e2e-waves $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e waves=$1 cw=$2 ms', round(23040e3/d['e2e']['value'],1))"; }
dev 1 1; dev 1 4; dev 6 1; dev 6 4; dev 6 8; dev 4 2; dev 8 2
e2e 32 8; e2e 48 4; e2e 24 4
