nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
PTX_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 6 --warmup 3 --no-e2e --no-secondary --no-cpu-baseline > gpurun_out/r2r_trace.json 2> gpurun_out/r2r_trace.err
grep -i "final\|trace" gpurun_out/r2r_trace.err | tail -40
tail -c 900 gpurun_out/r2r_trace.json
