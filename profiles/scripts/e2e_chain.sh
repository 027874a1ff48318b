# end-to-end (host buffers) step of the bench batch: queue kernel (cw 0 = auto picks it for 1152 chains) against the
# speculative kernel forced with cw warps per chain, at several wave counts
for cfg in ${CFGS:-"0 32" "2 32" "4 32" "2 16"}; do set -- $cfg; echo "cw=$1 e2e_waves=$2"; python bench.py --no-cpu-baseline --steps 2 --warmup 3 --chain-warps $1 --e2e-waves $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'], 23040e3/d['e2e']['value'], 'ms')"; done
