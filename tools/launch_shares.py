"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel: count, total us, share.
usage: launch_shares.py file.csv [last N launches]"""
import csv, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r["Kernel Name"], v))
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = k.split("(")[0][:90]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in rows)
print("%d launches, %.1f us in kernels" % (len(rows), tot))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.1f us %5.1f %%  x%-3d %s" % (v, 100 * v / tot, c, k))
