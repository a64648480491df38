timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in "" "PTX_NO_FILTER=1"; do echo "== $v"; env $v timeout 150 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print(j['value']/1e9, j['ms_per_step'], r['kernel_ms'], r['k_apply']['kernel_ms'], r['finalize_ms'], r['frac'], j['e2e']['ms_per_step'], j['parity'])
    else: print(l.rstrip())"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 3 -c 1 -o gpurun_out/r2d_apply python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r2d_ncu.log 2>&1; tail -1 gpurun_out/r2d_ncu.log | cut -c1-200
