"""Compact summary of `ncu --set full` reports: python tools/ncu_brief.py a.ncu-rep [b.ncu-rep ...]  (markdown table)"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__cluster_dim_x", "cluster"),
        ("launch__registers_per_thread", "regs"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
        ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM written MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed", "TMEM ld %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
rows = []
for f in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, vals = r[0], r[2]
    d = dict(zip(hdr, vals))
    rows.append((f.split("/")[-1].replace(".ncu-rep", ""), d))
print("| metric | " + " | ".join(n for n, _ in rows) + " |")
print("|---|" + "---|" * len(rows))
for k, label in KEYS:
    print(f"| {label} (`{k}`) | " + " | ".join(d.get(k, "") for _, d in rows) + " |")
