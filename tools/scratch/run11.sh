timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in "" "PTX_OLD_INGEST=1"; do echo "== c2 $v"; env $v timeout 300 python tools/bench_scatter.py c2 0 10 2>&1 | tail -1; done
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 3000 gpurun_out/r2f_bench.json
