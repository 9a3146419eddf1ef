#!/usr/bin/env python
"""Randomised GPU-vs-oracle parity sweep (test infrastructure, like tests/: the oracle is only the checker).
  tools/fuzz_parity.py [--seconds 240] [--seed 1] [--out gpurun_out/fuzz.json]
Every iteration draws a profile set (widths 1..1,600, amino or DNA), targets (ragged lengths, planted fragments, six-frame
reads or proteins), and pipeline options (F1/F2/F3, --nobias, --nonull2), runs the GPU search and the oracle, and compares the
hit set, coordinates, alignment and PP strings, scores and the filter funnel exactly as tests/test_gpu_search.py does."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from phyloaln_b200 import capi, hmmbuild_lite as hb, synth  # noqa: E402
from tests import helpers  # noqa: E402
from tests.test_gpu_search import compare_hits  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--out", default=None)
a = ap.parse_args()
t_end = time.time() + a.seconds
it, total_hits, log = 0, 0, []
while time.time() < t_end:
    seed = a.seed * 100003 + it
    rng = np.random.default_rng(seed)
    abc = "dna" if rng.random() < 0.2 else "amino"
    nh = int(rng.integers(1, 7))
    widths = [int(x) for x in np.clip(np.round(np.exp(rng.normal(np.log(150), 1.0, nh))), 1, 1600)]
    hmms = [helpers.calibrated_hmm(f"f{seed}_{i}", M, seed=seed * 7 + i, bg_mix=float(rng.choice([0.1, 0.25])), abc=abc, n=24)
            for i, M in enumerate(widths)]
    opts = {}
    if rng.random() < 0.3:
        opts["F1"] = float(rng.choice([0.002, 0.1, 0.5, 1.0]))
    if rng.random() < 0.3:
        opts["F2"] = float(rng.choice([1e-5, 1e-2, 0.3, 1.0]))
    if rng.random() < 0.3:
        opts["F3"] = float(rng.choice([1e-8, 1e-3, 1.0]))
    if rng.random() < 0.15:
        opts["do_bias"] = 0
    if rng.random() < 0.15:
        opts["do_null2"] = 0
    ctx = capi.Context(0)
    ctx.set_opts(F1=opts.get("F1", 0.02), F2=opts.get("F2", 1e-3), F3=opts.get("F3", 1e-5), do_bias=bool(opts.get("do_bias", 1)),
                 do_null2=bool(opts.get("do_null2", 1)))
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    mode = "dna" if abc == "dna" else ("reads" if rng.random() < 0.4 else "proteins")
    if mode == "reads":
        n = int(rng.integers(50, 400))
        reads = synth.random_reads(n, int(rng.integers(30, 260)), rng)
        synth.plant_reads(reads, hmms, 0.2, rng, sub_rate=0.03)
        rl = [r.tobytes().decode() for r in reads]
        ctx.load_read_list(rl, moltype=0, gencode=1)
        targets = [p for r in rl for _, p in O.six_frames(r, 1)]
    elif mode == "proteins":
        Lmax = int(rng.choice([40, 120, 600, 1800]))
        targets = helpers.protein_targets(hmms, int(rng.integers(50, 600 if Lmax < 600 else 120)), rng, Lmin=1, Lmax=Lmax,
                                          plant_every=int(rng.integers(2, 9)))
        ctx.load_proteins(targets)
    else:
        n = int(rng.integers(50, 300))
        reads = synth.random_reads(n, int(rng.integers(40, 300)), rng)
        for r in rng.choice(n, size=max(1, n // 5), replace=False):
            h = hmms[int(rng.integers(nh))]
            if h.M > 12:
                s0 = int(rng.integers(1, max(2, h.M - reads.shape[1])))
                frag = synth.sample_protein(h, rng, s0, min(h.M, s0 + reads.shape[1] - 1)).encode()
                reads[r, :len(frag)] = np.frombuffer(frag, dtype=np.uint8)
        rl = [r.tobytes().decode() for r in reads]
        ctx.load_read_list(rl, moltype=1)
        targets = [x for r in rl for x in (r, O.revcomp(r))]
    hits, model, aligned = ctx.search()
    oo = O.default_opts(**opts)
    nhit = compare_hits(O, hmms, targets, hits, model, aligned, opts=oo, pp=ctx.hit_pp(hits))
    st = ctx.stats()
    want = np.zeros(6, dtype=np.int64)
    for h in hmms:
        _, tr = O.Profile.from_hmm(h).search(targets, opts=oo, nthreads=os.cpu_count() or 4)
        want += np.bincount(tr["stage"], minlength=6)
    got = (st["n_past_msv"], st["n_past_bias"], st["n_past_vit"], st["n_past_fwd"])
    assert got == (want[1:].sum(), want[2:].sum(), want[3:].sum(), want[4:].sum()), (seed, got, want)
    ctx.close()
    total_hits += nhit
    log.append({"seed": seed, "abc": abc, "mode": mode, "widths": widths, "opts": opts, "targets": len(targets), "hits": nhit})
    it += 1
    print(f"iteration {it}: seed {seed} {mode} widths {widths} opts {opts}: {nhit} domains identical", flush=True)
res = {"iterations": it, "domains_compared": total_hits, "seconds": a.seconds, "cases": log}
print(json.dumps({k: res[k] for k in ("iterations", "domains_compared")}))
if a.out:
    json.dump(res, open(a.out, "w"), indent=0)
