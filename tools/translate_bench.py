#!/usr/bin/env python
"""Translation kernel throughput on a resident 1 Mi-read chunk (run on the B200 box)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyloaln_b200 import capi, synth
rng = np.random.default_rng(3)
h = synth.random_hmm("x", 100, rng)
ctx = capi.Context(0)
ctx.add_hmm(h, require_stats=False)
ctx.commit()
for n, nfrac in ((1 << 20, 0.001), (1 << 20, 0.0), (1 << 17, 0.001)):
    reads = synth.random_reads(n, 150, rng, n_frac=nfrac)
    offs = np.arange(n + 1, dtype=np.uint64) * 150
    best = 1e9
    for _ in range(4):
        ctx.load_reads(reads.reshape(-1), offs)
        best = min(best, ctx.stats()["ms_translate"])
    print(f"reads={n} N-frac={nfrac}: {best:.3f} ms  {446.0*n/(best*1e-3)/1e9:.0f} GB/s algorithmic (446 B/read)")
