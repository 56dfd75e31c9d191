"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_by_line.py src.csv [top_n]
Prints, per file, the lines with the most executed warp instructions and stall samples.
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, newline="")))
    fname = None
    hdr = None
    per = defaultdict(lambda: [0, 0, ""])   # (file, line) -> [instructions, samples, text]
    tot_i = tot_s = 0
    for r in rows:
        if len(r) >= 2 and r[0] in ("File Name", "File Path"):
            fname = r[1]; continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r; i_inst = hdr.index("Instructions Executed"); i_smp = hdr.index("# Samples"); continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "":   # a source line row (aggregated over its SASS)
            try:
                ins = int(r[i_inst]); smp = int(r[i_smp])
            except ValueError:
                continue
            key = (fname, int(r[0]))
            per[key][0] += ins; per[key][1] += smp; per[key][2] = r[1]
            tot_i += ins; tot_s += smp
    print(f"total warp instructions {tot_i:.4g}, samples {tot_s}")
    print("--- by instructions")
    for (f, ln), (ins, smp, txt) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * ins / max(tot_i, 1):5.1f}% inst {100 * smp / max(tot_s, 1):5.1f}% smp  {str(f).split('/')[-1]}:{ln}: {txt.strip()[:110]}")
    print("--- by stall samples")
    for (f, ln), (ins, smp, txt) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * ins / max(tot_i, 1):5.1f}% inst {100 * smp / max(tot_s, 1):5.1f}% smp  {str(f).split('/')[-1]}:{ln}: {txt.strip()[:110]}")


if __name__ == "__main__":
    main()
