"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report
(`ncu -i REP --page source --csv --print-source cuda,sass`).  Usage: ncu_lines.py REP [min_pct]"""
import csv, subprocess, sys, collections

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None
inst = collections.Counter(); samp = collections.Counter(); text = {}
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; hdr = None; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; i_n = hdr.index("# Samples"); i_ex = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= i_ex: continue
    if r[0]: cur = (fname, int(r[0])); text[cur] = r[1].strip()
    if r[2] and cur:  # a SASS row under the current source line
        num = lambda x: int(x) if x.isdigit() else 0
        inst[cur] += num(r[i_ex]); samp[cur] += num(r[i_n])
ti = sum(inst.values()); ts = sum(samp.values())
print(f"total warp instructions {ti}, samples {ts}")
for k in sorted(inst):
    pi = 100 * inst[k] / ti; ps = 100 * samp[k] / max(ts, 1)
    if pi >= min_pct or ps >= min_pct:
        print(f"{k[0]}:{k[1]:<5} inst {pi:5.2f}%  samples {ps:5.2f}%  {text[k][:110]}")
