"""Stall samples of one kernel by phase, from `ncu -i X.ncu-rep --page source --csv`: the instruction stream is cut at every
barrier / call / return / exit / mbarrier wait and the warp-stall samples of each piece are summed.
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv; python tools/stall_phases.py /tmp/src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
print(rows[0][1])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(f"total samples {tot}, {len(data)} instructions")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
cur, start, seg = 0, 0, collections.Counter()
for n, r in enumerate(data):
    cur += int(r[ix["# Samples"]] or 0)
    for h in stall_cols:
        if r[ix[h]]:
            seg[h] += int(r[ix[h]])
    src = r[ix["Source"]]
    if any(k in src for k in ("BAR.SYNC", "EXIT", "RET", "CALL", "SYNCS.PHASECHK.TRANS64.TRYWAIT")) or n == len(data) - 1:
        top = ", ".join(f"{k[6:]}={v}" for k, v in seg.most_common(3))
        if cur:
            print(f"instr {start:4d}-{n:4d}: {cur:6d} samples ({100 * cur / tot:5.1f} %)  ends at {src.strip()[:44]:44s} | {top}")
        cur, start, seg = 0, n + 1, collections.Counter()
