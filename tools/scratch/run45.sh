nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
for v in "X=1" "PTX_NO_SORT=1" "X=2" "PTX_NO_SORT=1"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 30 2>&1 | tail -1 | cut -c100-250; done
for v in "X=1" "PTX_NO_SORT=1"; do echo "== n50m $v"; env $v timeout 300 python tools/bench_scatter.py n50m 0 10 2>&1 | tail -1 | cut -c100-250; done
