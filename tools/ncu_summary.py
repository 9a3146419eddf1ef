#!/usr/bin/env python
"""Summarise an ncu raw CSV page: tools/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, re, io
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg"]
for r in rows[2:]:
    if not r:
        continue
    name = r[hdr.index("Kernel Name")]
    if pat and not pat.search(name):
        continue
    print(f"## {name}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
    for k in KEYS:
        if k in hdr:
            print(f"  {k:85s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_not_issued.ratio") is False and h.endswith(".ratio") and r[i]]
    for v, h in sorted(st, reverse=True)[:10]:
        print(f"  stall {h.split('issue_stalled_')[-1]:60s} {v:.3f}")
