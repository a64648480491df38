nvidia-smi --query-gpu=serial,pci.bus_id --format=csv,noheader
for lib in v_l256 v_l128 v_l128b v_l512; do
  r=$(PANTAX_GPU_LIB=$PWD/tools/scratch/libs/$lib.so timeout 300 python tools/bench_scatter.py c2 0 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ingest_ms'], d['apply_ms'], d['ms_per_step'], d['bases_checksum_first3'])")
  echo "$lib: $r"
done
