nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 100 --warmup 5 --no-e2e --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('N=2 value',d['value'],'ms/step',d['ms_per_step'],'ingest',r['kernel_ms'],'apply',r['k_apply']['kernel_ms'],'final',r['finalize_ms'],d['parity']['equal'])
"
