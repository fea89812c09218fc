python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9)"
