for v in "PTX_L2_HINTS=0" "PTX_L2_HINTS=1" "PTX_L2_HINTS=2" "PTX_L2_HINTS=3"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1 | cut -c1-330; done
for v in "PTX_L2_HINTS=0" "PTX_L2_HINTS=3"; do echo "== n50m $v"; env $v timeout 300 python tools/bench_scatter.py n50m 0 5 2>&1 | tail -1 | cut -c1-330; done
