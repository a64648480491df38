export PTX_L2_HINTS=1
for pad in 0 3072; do
PANTAX_GPU_LIB=$PWD/tools/scratch/libs/v_b8.so PTX_SMEM_PAD=$pad timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_s -s 4 -c 1 -o gpurun_out/r2h_ingest_b8_pad$pad python tools/bench_scatter.py c1 0 1 > gpurun_out/r2h_ncu_$pad.log 2>&1; tail -1 gpurun_out/r2h_ncu_$pad.log | cut -c1-200
done
