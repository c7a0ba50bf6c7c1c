"""Stall reasons per CUDA source line from an `ncu --page source --csv --print-source cuda,sass` dump:
    python tools/ncu_source_stalls.py dump.csv [top_n]
prints, for the hottest lines, the sampled stall reasons (all samples) as percentages of the line's samples."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, hdr = None, None
agg = defaultdict(lambda: defaultdict(int))
src = {}
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
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        i_samp = hdr.index("# Samples")
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    src[key] = r[1].strip()[:90]
    for i, name in stall_cols:
        try:
            agg[key][name] += int(r[i])
        except (ValueError, IndexError):
            pass
    try:
        agg[key]["_n"] += int(r[i_samp])
    except ValueError:
        pass
tot = sum(a["_n"] for a in agg.values())
allst = defaultdict(int)
for a in agg.values():
    for k, v in a.items():
        if k != "_n":
            allst[k] += v
print("all lines:", ", ".join(f"{k} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["_n"])[:top]:
    n = max(a["_n"], 1)
    st = ", ".join(f"{k} {100 * v / n:.0f}%" for k, v in sorted(((k, v) for k, v in a.items() if k != "_n"), key=lambda kv: -kv[1])[:3])
    print(f"{key[0]}:{key[1]:<5d} {100 * a['_n'] / max(tot, 1):5.1f}%  [{st}]  {src[key]}")
