nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 600 gpurun_out/r2q_bench.json; tail -3 gpurun_out/r2q_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2q_bench_reference.json 2>> gpurun_out/r2q_bench.err; tail -c 400 gpurun_out/r2q_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r2q_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_s -s 4 -c 1 -o gpurun_out/r2q_ingest_s python tools/bench_scatter.py c1 0 1 > gpurun_out/r2q_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 4 -c 1 -o gpurun_out/r2q_apply python tools/bench_scatter.py c1 0 1 > gpurun_out/r2q_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_l -s 3 -c 1 -o gpurun_out/r2q_ingest_l python tools/bench_scatter.py c2 0 1 > gpurun_out/r2q_ncu3.log 2>&1
ls -la gpurun_out/r2q*
