nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_driverlike.json 2> gpurun_out/r2z_bench.err; tail -c 300 gpurun_out/r2z_bench_driverlike.json; echo
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2>> gpurun_out/r2z_bench.err; tail -c 300 gpurun_out/r2z_bench.json; echo; tail -2 gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2>> gpurun_out/r2z_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r2z_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ingest_s -c 3 --csv --log-file gpurun_out/r2z_ingest_traffic.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/r2z_b2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_s -s 4 -c 1 -o gpurun_out/r2z_ingest_s python tools/bench_scatter.py c1 0 1 > gpurun_out/r2z_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 4 -c 1 -o gpurun_out/r2z_apply python tools/bench_scatter.py c1 0 1 > gpurun_out/r2z_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_l -s 3 -c 1 -o gpurun_out/r2z_ingest_l python tools/bench_scatter.py c2 0 1 > gpurun_out/r2z_ncu3.log 2>&1
python tools/bench_gfa.py 5000000 20
ls gpurun_out/r2z* | wc -l
