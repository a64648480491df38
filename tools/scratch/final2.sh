nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 900 python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; tail -c 200 gpurun_out/r2zz_bench.json; echo; tail -2 gpurun_out/r2zz_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2zz_bench_steps20.json 2>> gpurun_out/r2zz_bench.err
