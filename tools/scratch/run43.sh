timeout 600 python -m pytest tests/test_gpu_gfa.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_gfa.py 1000000 50; python tools/bench_gfa.py 5000000 20
