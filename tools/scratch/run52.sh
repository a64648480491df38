timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_scatter.py c1 0 30 2>&1 | tail -1 | cut -c100-330
timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2zzz_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2zzz_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'ingest',r['kernel_ms'],'apply',r['k_apply']['kernel_ms'],'final',r['finalize_ms'],'e2e',d['e2e']['value'],d['parity'])"
