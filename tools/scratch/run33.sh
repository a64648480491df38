nvidia-smi --query-gpu=index,name,temperature.gpu,temperature.memory,power.draw,clocks.sm,clocks.mem,ecc.mode.current --format=csv
for v in "X=1" "PTX_PF_WAVES=0" "PTX_PF_WAVES=3" "PTX_L2_HINTS=0" "PTX_L2_HINTS=3" "CUDA_VISIBLE_DEVICES=1" "X=2"; do echo "== c1 $v"; env $v timeout 300 python tools/bench_scatter.py c1 0 50 2>&1 | tail -1 | cut -c100-250; done
nvidia-smi --query-gpu=index,name,temperature.gpu,temperature.memory,power.draw,clocks.sm,clocks.mem --format=csv
