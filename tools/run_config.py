#!/usr/bin/env python
"""Run one BASELINE.json workload end to end at (a fraction of) its full size on one B200 and report
measured totals: tools/run_config.py --config {2,3,4} [--fraction 1.0] [--out gpurun_out/config2.json]

  config 2  10 M x 150 bp reads, six-frame, vs 1,000 BUSCO-size protein HMMs
  config 3  10 M x 150 bp reads, nucleotide targets (read + reverse complement) vs 1,000 DNA HMMs
            (M = 3 x the protein length distribution, clipped to this library's 2,047-column limit;
            PhyloAln's own --ref_split_len covers longer references)
  config 4  200 k transcripts of 0.5-5 kb (30 % carry one or two planted domains), six-frame, vs 3,000 protein HMMs
  config 5  100 M x 150 bp reads, six-frame, vs 3,000 protein HMMs (the per-GPU share of the 1/2/4/8-GPU job: reads shard by
            chunk, so one GPU at fraction f measures the per-read cost of the whole configuration)
Reads are generated chunk by chunk (seeded), uploaded from host memory, searched, and the hit records
pulled back: the timed region is the whole stream (generation excluded)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from phyloaln_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True)
ap.add_argument("--fraction", type=float, default=1.0)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rng = np.random.default_rng(3000 + a.config)
if a.config in (2, 3):
    n_units, n_hmm = int(10_000_000 * a.fraction), 1000
elif a.config == 5:
    n_units, n_hmm = int(100_000_000 * a.fraction), 3000
else:
    n_units, n_hmm = int(200_000 * a.fraction), 3000
abc = "dna" if a.config == 3 else "amino"
Ms = synth.busco_lengths(n_hmm, rng)
if a.config == 3:
    Ms = np.minimum(3 * Ms, 2047)
hmms = [synth.random_hmm(f"c{a.config}_{i:04d}", int(Ms[i]), rng, abc=abc) for i in range(n_hmm)]
ctx = capi.Context(0)
for h in hmms:
    ctx.add_hmm(h, require_stats=False)
ctx.commit()
t0 = time.perf_counter()
bench.calibrate_on_gpu(ctx, hmms, abc=abc)
t_cal = time.perf_counter() - t0
sumM = int(sum(h.M for h in hmms))
chunk = a.chunk or (131072 if a.config in (2, 3, 5) else 65536)
tot_s, tot_units, tot_hits, planted = 0.0, 0, 0, 0
funnel = {k: 0 for k in ("n_pairs", "n_cells", "n_past_ssv", "n_past_msv", "n_past_bias", "n_past_vit", "n_past_fwd", "n_domains")}
stage = {k: 0.0 for k in ("ms_translate", "ms_msv", "ms_bias", "ms_vit", "ms_fwd", "ms_domain", "ms_total")}
done = 0
vit_exact = 0
while done < n_units:
    n = min(chunk, n_units - done)
    if a.config in (2, 3, 5):
        reads = synth.random_reads(n, 150, rng)
        if abc == "amino":
            planted += len(synth.plant_reads(reads, hmms, 0.01, rng))
        else:  # plant fragments sampled from the DNA profiles themselves
            for r in rng.choice(n, size=n // 100, replace=False):
                h = hmms[int(rng.integers(n_hmm))]
                s = int(rng.integers(1, h.M - 150 + 1)) if h.M > 151 else 1
                frag = synth.sample_protein(h, rng, s, min(h.M, s + 149)).encode()
                reads[r, :len(frag)] = np.frombuffer(frag, dtype=np.uint8)
                planted += 1
        nt, off = reads.reshape(-1), np.arange(n + 1, dtype=np.uint64) * 150
    else:
        lens = rng.integers(500, 5001, n)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        nt = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(off[-1]))].copy()
        for r in np.nonzero(rng.random(n) < 0.3)[0]:
            for _ in range(int(rng.integers(1, 3))):
                h = hmms[int(rng.integers(n_hmm))]
                x = int(rng.integers(1, max(2, h.M // 2)))
                y = int(rng.integers(min(h.M, x + 30), h.M + 1))
                frag = np.frombuffer(synth.back_translate(synth.sample_protein(h, rng, x, y)).encode(), dtype=np.uint8)
                if len(frag) < lens[r]:
                    p = int(off[r]) + int(rng.integers(0, lens[r] - len(frag)))
                    nt[p:p + len(frag)] = frag
                    planted += 1
    t0 = time.perf_counter()
    ctx.load_reads(nt, off, moltype=1 if abc == "dna" else 0)
    hits, _, _ = ctx.search()
    tot_s += time.perf_counter() - t0
    st = ctx.stats()
    for k in funnel:
        funnel[k] += int(st[k])
    for k in stage:
        stage[k] += float(st[k])
    vit_exact += int(st["vit_exact_pairs"])
    tot_hits += len(hits)
    tot_units += n
    done += n
    print(f"  {done}/{n_units} units, {tot_units / tot_s:,.0f} units/s, domains so far {tot_hits}", file=sys.stderr, flush=True)
res = {"config": a.config, "fraction": a.fraction, "units": tot_units, "unit": "reads" if a.config != 4 else "transcripts",
       "n_hmm": n_hmm, "sum_M": sumM, "alphabet": abc, "seconds": tot_s, "units_per_s": tot_units / tot_s,
       "gcups": funnel["n_cells"] / tot_s / 1e9, "msv_gcups": funnel["n_cells"] / (stage["ms_msv"] * 1e-3) / 1e9,
       "planted": planted, "domains_reported_to_host": tot_hits, "funnel": funnel, "stage_ms": stage,
       "calibration_s": t_cal, "chunk": chunk, "vit_exact_pairs": vit_exact}
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
ctx.close()
