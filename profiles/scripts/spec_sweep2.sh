# segments per chain (32 x chain_warps) now independent of the block size: every CTA has >= 2 warps for the in-place fallback
dev() { python bench.py --no-cpu-baseline --no-e2e --waves $1 --chain-warps $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('device waves=$1 cw=$2 ms/step', round(d['ms_per_step'],2), 'chain ms', d['roofline']['kernel_ms_all']['k_att_chain'])"; }
e2e() { python bench.py --no-cpu-baseline --steps 2 --warmup 3 --waves 6 --chain-warps $2 --e2e-waves $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e waves=$1 cw=$2 ms', round(23040e3/d['e2e']['value'],1))"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
dev 1 1; dev 1 2; dev 6 1; dev 6 2
e2e 32 1; e2e 32 2
AME_CHAIN_WARPS=1 python profiles/scripts/single_track.py c2 | tail -2
