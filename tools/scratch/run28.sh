timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in "PTX_DS_CAS_FIRST=0" "PTX_DS_CAS_FIRST=1"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1 | cut -c100-330; done
timeout 300 python tools/bench_scatter.py c2 0 20 2>&1 | tail -1 | cut -c100-330
