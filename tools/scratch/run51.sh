PTX_STAGE_WALKS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in "PTX_STAGE_WALKS=0" "PTX_STAGE_WALKS=1"; do echo "== $v"; for s in c1 c2; do env $v timeout 300 python tools/bench_scatter.py $s 0 20 2>&1 | tail -1 | cut -c100-250; done; done
