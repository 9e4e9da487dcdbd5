"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import collections
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = collections.defaultdict(float)
cnt = collections.Counter()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
    name = row["Kernel Name"].split("(")[0]
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print(f"total kernel time {total/1e6:.2f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v/1e6:9.3f} ms  {100*v/total:5.1f}%  n={cnt[k]:5d}  avg={v/cnt[k]/1e3:8.1f} us  {k}")
