for s in c1 c2; do timeout 300 python tools/bench_scatter.py $s 0 20 2>&1 | tail -1 | cut -c1-400; done
