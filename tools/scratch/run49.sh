nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in c1 c2 n50m; do timeout 300 python tools/bench_scatter.py $s 0 20 2>&1 | tail -1 | cut -c100-330; done
