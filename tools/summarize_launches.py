"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/summarize_launches.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
i_name, i_val, i_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= i_val or r[i_metric] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[i_name]).replace("void kpd::", "").replace("kpd::", "")
    name = re.sub(r"<.*", "", name) + ("<" + ",".join(re.findall(r"\(int\)(\d+)", r[i_name])) + ">" if "<" in r[i_name] else "")
    a = agg[name]
    a[0] += 1
    a[1] += float(r[i_val].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:34s} n={a[0]:4d} total={a[1]:10.1f} us  avg={a[1] / a[0]:8.1f} us  share={100 * a[1] / tot:5.1f}%")
print(f"total {tot:.1f} us")
