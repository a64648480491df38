nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2z_bench_n8.json 2> gpurun_out/r2z_bench_n8.err; tail -c 600 gpurun_out/r2z_bench_n8.json; tail -3 gpurun_out/r2z_bench_n8.err
