nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
PANTAX_GPU_LIB=$PWD/tools/scratch/libs/v_dbg.so timeout 300 python tools/bench_scatter.py c1 0 2 2>&1 | grep -E "tile|shape" | tail -60 | cut -c1-200
