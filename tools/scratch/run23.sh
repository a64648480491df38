timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "PTX_LONG_NEW=1"; do echo "== c2 $v"; env $v timeout 300 python tools/bench_scatter.py c2 0 10 2>&1 | tail -1 | cut -c1-400; done
