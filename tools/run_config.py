#!/usr/bin/env python
"""Run one BASELINE.json workload end to end at (a fraction of) its full size on one B200 and report
measured totals: tools/run_config.py --config {2,3,4,5} [--fraction 1.0] [--out gpurun_out/config2.json]

  config 2  10 M x 150 bp reads, six-frame, vs 1,000 BUSCO-size protein HMMs
  config 3  10 M x 150 bp reads, nucleotide targets (read + reverse complement) vs 1,000 DNA HMMs
            (M = 3 x the protein length distribution, 300..6,000 columns; profiles beyond 2,047 run on csrc/wide.cuh)
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
ap.add_argument("--profile-region", action="store_true", help="cudaProfilerStart after calibration (ncu --profile-from-start off)")
a = ap.parse_args()
rng = synth.config_rng(a.config)
shape = synth.CONFIG_SHAPES[a.config]
n_units, n_hmm, abc = int(shape["units"] * a.fraction), shape["n_hmm"], shape["abc"]
hmms = synth.config_profiles(a.config, rng)
ctx = capi.Context(0)
for h in hmms:
    ctx.add_hmm(h, require_stats=False)
ctx.commit()
t0 = time.perf_counter()
bench.calibrate_on_gpu(ctx, hmms, abc=abc)
t_cal = time.perf_counter() - t0
sumM = int(sum(h.M for h in hmms))
chunk = a.chunk or shape["chunk"]
tot_s, tot_units, tot_hits, planted = 0.0, 0, 0, 0
funnel = {k: 0 for k in ("n_pairs", "n_cells", "n_past_ssv", "n_past_msv", "n_past_bias", "n_past_vit", "n_past_fwd", "n_domains")}
stage = {k: 0.0 for k in ("ms_translate", "ms_msv", "ms_bias", "ms_vit", "ms_fwd", "ms_domain", "ms_total")}
done = 0
vit_exact = 0
if a.profile_region:
    capi.profiler_range(True)
while done < n_units:
    n = min(chunk, n_units - done)
    nt, off, npl = synth.config_chunk(a.config, n, rng, hmms)
    planted += npl
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
       "n_hmm": n_hmm, "sum_M": sumM, "max_M": int(max(h.M for h in hmms)), "n_wide_profiles": int(sum(h.M > 2047 for h in hmms)), "alphabet": abc, "seconds": tot_s, "units_per_s": tot_units / tot_s,
       "gcups": funnel["n_cells"] / tot_s / 1e9, "msv_gcups": funnel["n_cells"] / (stage["ms_msv"] * 1e-3) / 1e9,
       "planted": planted, "domains_reported_to_host": tot_hits, "funnel": funnel, "stage_ms": stage,
       "calibration_s": t_cal, "chunk": chunk, "vit_exact_pairs": vit_exact}
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
ctx.close()
