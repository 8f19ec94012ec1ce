"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [launches_per_step]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for d in csv.DictReader(lines):
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    rows.append((re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", ""), v, d["Grid Size"]))
if per_step:
    rows = rows[:per_step]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, g in rows:
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v, _ in rows)
print(f"{len(rows)} launches, {tot:.1f} us total (cold-cache, serialised under ncu: compare shares, not absolutes)")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {v:.1f} | {100 * v / tot:.1f}% | {v / c:.1f} |")
