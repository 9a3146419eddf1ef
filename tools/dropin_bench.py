#!/usr/bin/env python
"""Whole drop-in step on files, as PhyloAln would run it: fq.gz -> native prepare_target_step -> hmmsearch_step (all
profiles x all reads on the GPU, hit_info.tsv / targets.fas / ok markers per alignment) with wall-clock per phase.
  tools/dropin_bench.py [--reads 1000000] [--hmms 200] [--out gpurun_out/dropin.json]
Also times the same search entered from the reference's own all.raw.fasta (native pa_fasta loader)."""
import argparse
import gzip
import json
import os
import shutil
import sys
import tempfile
import time
from argparse import Namespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from phyloaln_b200 import capi, dropin, hmmio, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--hmms", type=int, default=200)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rng = np.random.default_rng(77)
Ms = synth.busco_lengths(a.hmms, rng)
hmms = [synth.random_hmm(f"OG{i:05d}", int(Ms[i]), rng) for i in range(a.hmms)]
ctx = capi.Context(0)
for h in hmms:
    ctx.add_hmm(h, require_stats=False)
ctx.commit()
bench.calibrate_on_gpu(ctx, hmms)
ctx.close()
tmp = tempfile.mkdtemp(prefix="pa_dropin_")
os.makedirs(os.path.join(tmp, "ref_hmm"))
for h in hmms:
    hmmio.write_hmm(h, os.path.join(tmp, "ref_hmm", h.name + ".hmm"))
reads = synth.random_reads(a.reads, 150, rng)
synth.plant_reads(reads, hmms, 0.01, rng)
t0 = time.perf_counter()
files = []
for part in range(2):   # paired files of one species
    p = os.path.join(tmp, f"SP_{part + 1}.fq.gz")
    with gzip.open(p, "wb", compresslevel=1) as f:
        half = reads[part::2]
        q = b"I" * 150
        f.write(b"".join(b"@r%d/%d\n" % (i, part + 1) + half[i].tobytes() + b"\n+\n" + q + b"\n" for i in range(len(half))))
    files.append(p)
t_gen = time.perf_counter() - t0
res = {"reads": a.reads, "n_hmm": a.hmms, "generate_fastq_s": t_gen}
args = Namespace(file_format="guess", split_len=None, split_slide=None, merge_len=250, cpu=os.cpu_count() or 1)
t0 = time.perf_counter()
ing = dropin.prepare_target_step(tmp, {"SP": files}, args)
res["prepare_target_step_s"] = time.perf_counter() - t0
groups = [h.name for h in hmms]
t0 = time.perf_counter()
counts = dropin.hmmsearch_step(tmp, groups, mol_type="prot", gencode=1, ingest=ing, write_text=False)
res["hmmsearch_step_from_ingest_s"] = time.perf_counter() - t0
res["phases_from_ingest"] = dict(dropin.hmmsearch_step.last_timing)
res["hit_rows"] = int(sum(counts.values()))
shutil.rmtree(os.path.join(tmp, "map"))
shutil.rmtree(os.path.join(tmp, "ok"))
t0 = time.perf_counter()
counts2 = dropin.hmmsearch_step(tmp, groups, mol_type="prot", gencode=1, write_text=False)   # reads all.raw.fasta
res["hmmsearch_step_from_raw_fasta_s"] = time.perf_counter() - t0
res["phases_from_raw_fasta"] = dict(dropin.hmmsearch_step.last_timing)
assert counts2 == counts
res["reads_per_s_whole_step"] = a.reads / (res["prepare_target_step_s"] + res["hmmsearch_step_from_ingest_s"])
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
shutil.rmtree(tmp, ignore_errors=True)
