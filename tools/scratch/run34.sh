nvidia-smi | head -20
nvidia-smi -q -i 0 | grep -i -E "driver|vbios|Product|Memory|bus id|Max|Persistence|MIG|Clocks Event|Fabric|C2C|Confidential|Compute Mode|Addressing|Serial" | head -60
lscpu | grep -E "Model name|Socket|^CPU\(s\)|NUMA" | head
python - <<'PY'
import torch,time
x=torch.empty(2_000_000_000,dtype=torch.uint8,device='cuda'); y=torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): y.copy_(x)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print("copy GB/s (r+w)", 2*x.numel()/dt/1e9)
PY
timeout 300 python tools/bench_scatter.py c1 0 50 2>&1 | tail -1 | cut -c100-250
