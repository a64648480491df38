export PTX_L2_HINTS=1
for rep in 1 2; do
for lib in head v_b8 v_b7 v_b7s16 v_b6; do
for pad in 0 3072 8192; do
  r=$(PANTAX_GPU_LIB=$PWD/tools/scratch/libs/$lib.so PTX_SMEM_PAD=$pad timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ingest_ms'], d['apply_ms'], d['ms_per_step'])")
  echo "$lib pad=$pad: $r"
done; done; done
