for i in 1 2 3 4 5 6 7 8; do timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ingest_ms'], d['apply_ms'], d['ms_per_step'])"; done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
