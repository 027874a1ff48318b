# device-resident step against the tile sizes of the time-parallel filter kernels (0 = the plan's own choice)
run() { python bench.py --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_all']; print('$*', '->', round(d['ms_per_step'],2), 'ms/step  eq', k['k_eq'], 'split', k['k_band_split'], 'kw', k['k_kweight_energy'])"; }
run --kw-tile 0
run --kw-tile 3
run --kw-tile 4
run --kw-tile 6
run --eq-tile 32768
run --eq-tile 12288
run --xover-tile 16384
run --xover-tile 6144
