"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: python tools/launch_summary.py list.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    name, t = r[4].split("(")[0], float(r[-1])
    agg[name][0] += 1
    agg[name][1] += t
    agg[name][2] = max(agg[name][2], t)
total = sum(v[1] for v in agg.values())
for name, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:40s} {v[0]:6d} launches {v[1] / 1e3:11.1f} us {100 * v[1] / total:5.1f} %   avg {v[1] / v[0] / 1e3:8.2f} us   max {v[2] / 1e3:8.1f} us")
print(f"total {total / 1e6:.3f} ms in {len(rows)} launches")
