export PTX_L2_HINTS=1
for v in "X=0" "PTX_CLS_MUL=1" "PTX_SMEM_PAD=3072" "PTX_CLS_MUL=1 PTX_SMEM_PAD=3072"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1 | cut -c100-330; done
