timeout 420 python -m pytest tests -m gpu -x -q -k "scatter_variant or kats or short_reads_multi" 2>&1 | tail -6
for sh in c1 c2 n50m hot; do for v in 0 1 2 3; do timeout 300 python tools/bench_scatter.py $sh $v 10 2>&1 | tail -1; done; done | tee gpurun_out/r2_scatter_times.jsonl
for sh in c1 c2 n50m; do for v in 0 1 2 3; do
 timeout 400 ncu --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_apply|DeviceRadixSort|DeviceReduce|k_add_runs" -s $([ $v = 3 ] && echo 0 || echo 3) -c $([ $v = 3 ] && echo 400 || echo 1) --csv --log-file gpurun_out/r2_scatter_ncu_${sh}_v${v}.csv python tools/bench_scatter.py $sh $v 1 > /dev/null 2>&1
done; done
ls gpurun_out | grep r2_scatter | head -30
