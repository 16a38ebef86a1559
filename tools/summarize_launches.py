"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, share."""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9}.get(unit, 1)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    rows.append((int(r["ID"]), name, ns, r["Grid Size"], r["Block Size"]))
rows = [r for r in rows if r[0] >= skip]
tot = sum(r[2] for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for _, n, ns, *_ in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"launches: {len(rows)}  total device time: {tot / 1e6:.3f} ms (cold-cache, serialised under ncu: compare shares)")
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n[:90]}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")
