nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "X=1" "PTX_LONG_TILE=32768" "PTX_SMEM_PAD_L=8192" "PTX_LONG_TILE=24576" "PTX_LONG_TILE=20480" "PTX_LONG_TILE=16384"; do echo "== c2 $v"; env $v timeout 300 python tools/bench_scatter.py c2 0 20 2>&1 | tail -1 | cut -c100-330; done
