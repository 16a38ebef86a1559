"""DRAM traffic of ONE timestep from an ncu launch list taken with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` over `bench.py`.

The launches between the last two `step_vpsde_kernel` launches (= one full timestep: M score-net forwards + the fused
step) are grouped per kernel: launches, time, DRAM read / written.  Writes a JSON summary (second argument) that `bench.py`
quotes as `roofline.traffic`.

    python tools/summarize_traffic.py gpurun_out/traffic.csv profiles/r01d_step_traffic.json
"""
import collections
import csv
import json
import re
import sys

path = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
with open(path) as fh:
    lines = [l for l in fh if l.startswith('"')]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6,
        "msecond": 1e6}
launch = collections.OrderedDict()
for r in csv.DictReader(lines):
    i = int(r["ID"])
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
    d = launch.setdefault(i, {"name": name})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1)
ids = sorted(launch)
marks = [i for i in ids if "step_vpsde_kernel" in launch[i]["name"]]
if len(marks) < 2:
    sys.exit("need at least two step_vpsde_kernel launches in the list")
lo, hi = marks[-2], marks[-1]
step = [launch[i] for i in ids if lo < i <= hi]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in step:
    a = agg[d["name"]]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot_ns = sum(a[1] for a in agg.values())
print(f"one timestep = launches {lo + 1}..{hi}: {len(step)} launches, {tot_ns / 1e6:.3f} ms serialised under ncu")
print("| kernel | launches | ms | DRAM read MB | DRAM written MB |")
print("|---|---:|---:|---:|---:|")
summary = {"launches": len(step), "kernels": {}}
for n, (c, ns, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n[:70]}` | {c} | {ns / 1e6:.3f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} |")
    summary["kernels"][n[:70]] = {"launches": c, "ms": ns / 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr}
tens = [v for k, v in summary["kernels"].items() if "gemm_tcgen05" in k or "attn_core" in k]
summary["tensor_core_kernels"] = {"launches": sum(v["launches"] for v in tens),
                                  "dram_bytes": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in tens)}
summary["all_kernels_dram_bytes"] = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in summary["kernels"].values())
print(f"tensor-core kernels: {summary['tensor_core_kernels']['launches']} launches, "
      f"{summary['tensor_core_kernels']['dram_bytes'] / 1e9:.3f} GB; whole timestep {summary['all_kernels_dram_bytes'] / 1e9:.3f} GB")
if out:
    with open(out, "w") as fh:
        json.dump(summary, fh, indent=1)
