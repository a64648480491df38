nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -5
