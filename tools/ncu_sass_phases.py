"""Per source line: executed warp instructions per site, samples, for a line range; from `ncu --page source --csv --print-source cuda,sass`.
    python tools/ncu_sass_phases.py src.csv n_sites file_substring lo hi
"""
import csv, sys
from collections import defaultdict
path, nsites, fsub, lo, hi = sys.argv[1], float(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
rows = list(csv.reader(open(path, newline="")))
fname = None; hdr = None; cur = None
tot = defaultdict(lambda: [0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; i_inst = hdr.index("Instructions Executed"); i_smp = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        try:
            ln = int(r[0]); ins = int(r[i_inst]); smp = int(r[i_smp])
        except ValueError:
            continue
        if fsub in fname and lo <= ln <= hi:
            tot[ln][0] += ins; tot[ln][1] += smp; tot[ln][2] = r[1]
ti = sum(v[0] for v in tot.values()); ts = sum(v[1] for v in tot.values())
print(f"lines {lo}-{hi} of {fsub}: {ti / nsites:.1f} warp-instr/site, {ts} samples")
for ln in sorted(tot):
    v = tot[ln]
    if v[0] / nsites >= 2 or v[1] >= 50:
        print(f"{ln:5d} {v[0] / nsites:8.1f} i/site {v[1]:6d} smp  {v[2].strip()[:120]}")
