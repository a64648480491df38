for v in "" "PTX_L2_FETCH=32" "PTX_L2_FETCH=128"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 4 -c 1 -o gpurun_out/r2g_apply python tools/bench_scatter.py c1 0 1 > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log | cut -c1-300
