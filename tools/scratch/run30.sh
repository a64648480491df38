nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown --format=csv -lms 50 > gpurun_out/r2l_smi.csv 2>&1 &
SMI=$!
for v in "PTX_DS_CAS_FIRST=0" "PTX_DS_CAS_FIRST=1"; do echo "== c1 $v"; date +%T.%N; env $v timeout 300 python tools/bench_scatter.py c1 0 400 2>&1 | tail -1 | cut -c100-330; date +%T.%N; done
kill $SMI
awk -F, '{print $2, $4}' gpurun_out/r2l_smi.csv | sort | uniq -c | sort -k1 -n -r | head -20
