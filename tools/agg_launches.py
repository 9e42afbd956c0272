"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/agg_launches.py file.csv [--list]"""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
rows = []
for row in csv.DictReader(lines):
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    k = row["Kernel Name"][:64]
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
    rows.append((int(row["ID"]), k, row["Grid Size"], v))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} us {n:5d} {100 * t / tot:5.1f}% {k}")
print(f"{tot:10.1f} us total, {len(rows)} launches")
if "--list" in sys.argv:
    for r in rows:
        print(f"{r[0]:5d} {r[3]:8.1f} {r[2]:>16s} {r[1]}")
