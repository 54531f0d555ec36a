"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1
    agg[name][1] += us
    total += us
print(f"total {total/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1e3:9.3f} ms  {100*us/total:5.1f}%  n={n:4d}  avg {us/n:9.1f} us  {name}")
