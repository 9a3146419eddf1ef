#!/usr/bin/env python
"""Per-kernel time shares from an ncu launch list (--metrics gpu__time_duration.sum --csv): tools/launch_shares.py launches.csv"""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if r and r[0] == "ID":
        hdr = r
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    tot[name] += v
    cnt[name] += 1
all_ms = sum(tot.values())
print(f"total {all_ms:.1f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.2f} ms {100 * v / all_ms:6.2f} %  x{cnt[k]:<5d} {k}")
