#!/usr/bin/env python
"""Per-Q throughput of the SSV/MSV stage: tools/msv_sweep.py [--impl tmem|smem] (run on the B200 box).
Prints padded and useful cells/clk/SM per model length."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="tmem")
ap.add_argument("--Ms", default="60,120,250,300,380,450,500,560,700,1000")
ap.add_argument("--nh", type=int, default=48)
ap.add_argument("--reads", type=int, default=32768)
a = ap.parse_args()
os.environ["PA_MSV_IMPL"] = a.impl
from phyloaln_b200 import capi, synth

rng = np.random.default_rng(5)
reads = synth.random_reads(a.reads, 150, rng)
offs = np.arange(a.reads + 1, dtype=np.uint64) * 150
for M in [int(x) for x in a.Ms.split(",")]:
    ctx = capi.Context(0)
    hmms = [synth.random_hmm(f"m{M}_{i}", M, rng) for i in range(a.nh)]
    for h in hmms:
        ctx.add_hmm(h, require_stats=False)
    ctx.commit()
    ctx.load_reads(reads.reshape(-1), offs)
    best = 1e9
    for rep in range(3):
        ctx.search()
        best = min(best, ctx.stats()["ms_msv"])
    Q = M // 64 + 1
    cells = 296.0 * a.reads * a.nh
    clk = 1.965e9 * 148
    print(f"impl={a.impl} M={M:5d} Q={Q:2d} ms_msv={best:8.3f}  useful {cells*M/(best*1e-3)/clk:6.1f}  padded {cells*64*Q/(best*1e-3)/clk:6.1f} cells/clk/SM"
          f"  ({cells*M/(best*1e-3)/1e12:5.2f} T useful cells/s)", flush=True)
    ctx.close()
