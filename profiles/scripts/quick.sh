python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --no-cpu-baseline --no-e2e --waves 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1 wave', round(d['ms_per_step'],2), d['roofline']['kernel_ms_all'])"
python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('6 waves', round(d['ms_per_step'],2), round(d['value']))"
python profiles/scripts/single_track.py c2 | tail -2
