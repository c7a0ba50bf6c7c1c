"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
    python tools/ncu_source_hot.py dump.csv [top_n]
prints executed warp-instructions and stall samples per (file, line), hottest first."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    a = agg[(cur_file, line)]
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    a[0] += num(r[i_inst])
    a[1] += num(r[i_samp])
    a[2] = r[1].strip()[:110]
tot_i = sum(a[0] for a in agg.values())
tot_s = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:<5d} inst {100 * a[0] / max(tot_i, 1):5.1f}%  samples {100 * a[1] / max(tot_s, 1):5.1f}%  {a[2]}")
