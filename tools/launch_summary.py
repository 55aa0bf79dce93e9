"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) per kernel name.
   python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n_launches] > profiles/rNN_launches_summary.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    rows.append((r["Kernel Name"], us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for name, us in rows:
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    d = agg.setdefault(name, [0.0, 0])
    d[0] += us
    d[1] += 1
tot = sum(d[0] for d in agg.values())
print(f"# {len(rows)} launches, {tot:.1f} us of kernel time")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:9.1f} us {100 * us / tot:5.1f}% x{n:4d} avg {us / n:7.1f}  {name[:100]}")
