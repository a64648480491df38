nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_gfa.py tests/test_host_driver.py -m gpu -x -q 2>&1 | tail -15
