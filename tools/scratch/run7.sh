timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 200 python bench.py --steps 50 --warmup 3 --no-secondary 2>&1 | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print(j['value']/1e9, j['ms_per_step'], r['kernel_ms'], r['k_apply']['kernel_ms'], r['finalize_ms'], r['frac'], j['e2e'], j['parity'], j['cpu_baseline'])
    else: print(l.rstrip())"
