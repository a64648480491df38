nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 --secondary-records 20000000 2>&1 | tail -3 > gpurun_out/r2_bench_n2.log; cat gpurun_out/r2_bench_n2.log | cut -c1-4000
