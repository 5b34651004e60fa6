"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv [n_steps]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = rows[1:]
last = data[-(len(data) // steps):]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in last:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"void |at::native::|<unnamed>::", "", name)[:100]
    t = float(r[vi].replace(",", "")) / (1000.0 if r[ui] in ("ns", "nsecond") else 1.0)
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print("launches/step", len(last), "total us", round(tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1]:9.0f} us {100 * v[1] / tot:5.1f}% x{v[0]:4d}  {k}")
