"""Summarise an `ncu --set full` report of one kernel into profiles/<tag>_<kernel>_ncu_summary.md; for the dominant
kernel (k_ingest) also write profiles/r1_ingest_traffic.json (read by bench.py for roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep <tag> <records> [ingest|apply]
"""
import csv
import io
import json
import subprocess
import sys
import collections

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']


def main():
    rep, tag, records = sys.argv[1], sys.argv[2], int(sys.argv[3])
    kern = sys.argv[4] if len(sys.argv) > 4 else 'ingest'
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, d = rows[0], rows[1], rows[2]
    val = {n: (d[i], units[i]) for i, n in enumerate(h)}
    kname = {"ingest": "k_ingest (GAF parse -> labels, counts, record table, CSR walks)", "apply": "k_apply<CLASSIFY|COVER> (id set + node coverage from the CSR)"}[kern]
    out = [f"# {tag}: `ncu --set full --clock-control none --import-source on` of {kname}, config 2 ({records} records)\n",
           "| metric | value | unit |", "|---|---|---|"]
    for k in KEEP:
        if k in val:
            out.append(f"| {k} | {val[k][0]} | {val[k][1]} |")
    out.append("\nWarp stall reasons (warps per issue slot):\n")
    for n, (v, _u) in sorted(val.items()):
        if n.startswith('smsp__average_warp') and 'issue_stalled' in n and 'not_issued' not in n:
            try:
                if float(v) > 0.3:
                    out.append(f"* {n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {float(v):.2f}")
            except ValueError:
                pass
    sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(sass)))
    sh, data = srows[1], srows[2:]
    isrc, ie, it, ns = sh.index('Source'), sh.index('Instructions Executed'), sh.index('Thread Instructions Executed'), sh.index('# Samples')
    ops = collections.Counter()
    tot = 0
    for r in data:
        if not r[ie].isdigit():
            continue
        txt = r[isrc].strip()
        op = txt.split()[1] if txt.startswith('@') else txt.split()[0]
        ops[op.split('.')[0]] += int(r[ie])
        tot += int(r[ie])
    out.append(f"\nWarp-level instruction mix ({tot/1e6:.0f} M warp instructions; top opcodes):\n")
    for op, n in ops.most_common(14):
        out.append(f"* {op}: {n/1e6:.1f} M ({100*n/tot:.1f} %)")
    tma = sum(n for op, n in ops.items() if op in ('UBLKCP', 'UTMALDG', 'SYNCS'))
    out.append(f"\nTMA / mbarrier instructions present: {', '.join(op for op in ('UBLKCP', 'SYNCS', 'UTMALDG') if ops.get(op))} ({tma} executed)")
    dram = float(val['dram__bytes_read.sum'][0]) * (1e9 if val['dram__bytes_read.sum'][1] == 'Gbyte' else 1e6) + \
        float(val['dram__bytes_write.sum'][0]) * (1e9 if val['dram__bytes_write.sum'][1] == 'Gbyte' else 1e6)
    dur_ms = float(val['gpu__time_duration.sum'][0]) * (1.0 if val['gpu__time_duration.sum'][1] == 'ms' else 1e-3)
    atom = sum(float(val[k][0]) for k in ('l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum') if k in val)
    out.append(f"\nDerived: DRAM traffic {dram/1e9:.3f} GB per launch = {dram/1e9/(dur_ms*1e-3)/1e3:.2f} TB/s under ncu; "
               f"global atomic + reduction sectors {atom/1e6:.1f} M = {atom/(dur_ms*1e-3)/1e9:.1f} G sectors/s.")
    open(f"profiles/{tag}_{kern}_ncu_summary.md", "w").write("\n".join(out) + "\n")
    if kern == "ingest":
        json.dump({"records": records, "dram_bytes_per_launch": dram, "source": f"profiles/{tag}_ingest_ncu_summary.md",
                   "kernel": "k_ingest"}, open("profiles/r1_ingest_traffic.json", "w"))
    print("\n".join(out))


if __name__ == "__main__":
    main()
