"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us; tot += us
print(f"total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1e3:9.3f} ms  {100*us/tot:5.1f}%  n={c:6d}  avg={us/c:8.2f} us  {name[:110]}")
