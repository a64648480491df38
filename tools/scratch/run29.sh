timeout 1500 ncu --replay-mode application --cache-control none --clock-control none --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on -k regex:k_ingest_s -s 8 -c 1 -o gpurun_out/r2k_app_slow python tools/bench_scatter.py c1 0 3 > gpurun_out/r2k_app_slow.log 2>&1
tail -2 gpurun_out/r2k_app_slow.log | cut -c1-200
