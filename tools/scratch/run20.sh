for rep in 1 2; do
for lib in v_t128 v_t256b3 v_t256b4 v_t64; do
  r=$(PANTAX_GPU_LIB=$PWD/tools/scratch/libs/$lib.so timeout 300 python tools/bench_scatter.py c1 0 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ingest_ms'], d['apply_ms'], d['ms_per_step'], d['bases_checksum_first3'])")
  echo "$lib: $r"
done; done
