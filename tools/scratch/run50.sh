for v in "PTX_NI_PREFETCH=0" "PTX_NI_PREFETCH=3" "PTX_NI_PREFETCH=8"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1 | cut -c100-330; done
