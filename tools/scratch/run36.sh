nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
for lib in v_a8 v_a6 v_a5 v_a4; do
  r=$(PANTAX_GPU_LIB=$PWD/tools/scratch/libs/$lib.so timeout 300 python tools/bench_scatter.py c1 0 30 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ingest_ms'], d['apply_ms'], d['ms_per_step'], d['bases_checksum_first3'])")
  echo "$lib: $r"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest_s -s 3 -c 1 -o gpurun_out/r2p_ingest_ms python tools/bench_scatter.py n50m 0 1 > gpurun_out/r2p_ncu.log 2>&1; tail -1 gpurun_out/r2p_ncu.log | cut -c1-200
