for v in "PTX_PF_WAVES=0" "PTX_PF_WAVES=1" "PTX_PF_WAVES=2" "PTX_PF_WAVES=0" "PTX_PF_WAVES=1"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 50 2>&1 | tail -1 | cut -c100-330; done
for v in "PTX_PF_WAVES=0" "PTX_PF_WAVES=1"; do echo "== c2 $v"; env $v timeout 300 python tools/bench_scatter.py c2 0 20 2>&1 | tail -1 | cut -c100-330; done
