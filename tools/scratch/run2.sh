# every GPU command under its own timeout: a hung kernel must never run into gpurun's limit
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in "" "PTX_NO_SORT=1" "PTX_TILE_BYTES=12288" "PTX_TILE_BYTES=16384" "PTX_L2_FETCH=32" "PTX_L2_FETCH=128"; do echo "== $v"; env $v timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary 2>&1 | python -c "import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print(j['value']/1e9, j['ms_per_step'], r['kernel_ms'], r['k_apply']['kernel_ms'], r['finalize_ms'], r['frac'], j['clocks'])
    else: print(l.rstrip())"; done
