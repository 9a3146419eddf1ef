#!/usr/bin/env python
"""File-to-file measurement of the drop-in step as PhyloAln would run it, on 1..N GPUs of one box:

    fq.gz files -> prepare_target_step (native ingest, streaming) -> hmmsearch_step(devices=range(N))
               -> map/{g}.hit_info.tsv + map/{g}.targets.fas + ok/ markers for every alignment g

  tools/dropin_bench.py --config 2 [--reads 10000000] [--gpus 1,2,4,8] [--files 16] [--out gpurun_out/f2f.json]

--config 2: BASELINE.json config 2 (1,000 protein HMMs); --config 5: the 3,000-HMM set of config 5 (a slice of its 100 M
reads).  The timed region starts with the compressed FASTQ files on disk and ends when the last ok/ marker is written; the
ingest overlaps the search (wait=False).  Per-phase times, the GPU-resident search time summed over contexts and the
limiter (the slowest phase) are reported next to reads/s.  Input generation, profile calibration and writing ref_hmm/*.hmm
are outside the timed region (they are hmmbuild's and the sequencer's work)."""
import argparse
import gzip
import json
import multiprocessing as mp
import os
import shutil
import sys
import tempfile
import time
from argparse import Namespace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _write_fastq(job):
    """One gzip'ed FASTQ file of the config's reads (own seeded stream per file; planted fragments from the shared profiles)."""
    path, config, n, seed, hmm_dir, names = job
    from phyloaln_b200 import hmmio, synth
    rng = np.random.default_rng(seed)
    hmms = [hmmio.read_hmm(os.path.join(hmm_dir, g + ".hmm")) for g in names[:: max(1, len(names) // 64)]]   # planting donors
    q = b"I" * 150
    with gzip.open(path, "wb", compresslevel=1) as f:
        done = 0
        while done < n:
            m = min(65536, n - done)
            reads = synth.random_reads(m, 150, rng)
            synth.plant_reads(reads, hmms, 0.01, rng)
            tag = os.path.basename(path).split(".")[0].encode()
            f.write(b"".join(b"@" + tag + b".%d\n" % (done + i) + reads[i].tobytes() + b"\n+\n" + q + b"\n" for i in range(m)))
            done += m
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--hmms", type=int, default=0)
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--files", type=int, default=16)
    ap.add_argument("--chunk", type=int, default=1 << 18)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import bench
    from phyloaln_b200 import capi, dropin, hmmio, synth
    shape = synth.CONFIG_SHAPES[a.config]
    assert shape["abc"] == "amino" and shape["unit"] == "reads", "file-to-file runs use the read configs (2, 5)"
    rng = synth.config_rng(a.config)
    hmms = synth.config_profiles(a.config, rng, n_hmm=a.hmms or None)
    ctx = capi.Context(0)
    for h in hmms:
        ctx.add_hmm(h, require_stats=False)
    ctx.commit()
    bench.calibrate_on_gpu(ctx, hmms)
    ctx.close()
    tmp = tempfile.mkdtemp(prefix="pa_f2f_")
    os.makedirs(os.path.join(tmp, "ref_hmm"))
    names = [h.name for h in hmms]
    for h in hmms:
        hmmio.write_hmm(h, os.path.join(tmp, "ref_hmm", h.name + ".hmm"))
    t0 = time.perf_counter()
    nsp = max(1, a.files // 2)
    per = a.reads // a.files
    jobs, rawdata = [], {}
    for s in range(nsp):
        for part in range(2 if a.files > 1 else 1):
            p = os.path.join(tmp, f"SP{s:02d}_{part + 1}.fq.gz")
            jobs.append((p, a.config, per, 1000 + 2 * s + part, os.path.join(tmp, "ref_hmm"), names))
            rawdata.setdefault(f"SP{s:02d}", []).append(p)
    with mp.get_context("spawn").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        pool.map(_write_fastq, jobs)
    nreads = per * len(jobs)
    res = {"config": a.config, "workload": shape["what"], "reads": nreads, "n_hmm": len(hmms), "input_files": len(jobs),
           "input_bytes_gz": int(sum(os.path.getsize(j[0]) for j in jobs)), "generate_fastq_s": time.perf_counter() - t0,
           "host_cores": os.cpu_count(), "runs": []}
    args = Namespace(file_format="guess", split_len=None, split_slide=None, merge_len=250, cpu=os.cpu_count() or 1)
    for ngpu in [int(x) for x in a.gpus.split(",")]:
        out = os.path.join(tmp, f"run{ngpu}")
        os.makedirs(out)
        os.symlink(os.path.join(tmp, "ref_hmm"), os.path.join(out, "ref_hmm"))
        t0 = time.perf_counter()
        ing = dropin.prepare_target_step(out, rawdata, args, wait=False)
        t1 = time.perf_counter()
        t_ing = [None]

        def watch():
            while not ing.done():
                time.sleep(0.005)
            t_ing[0] = time.perf_counter() - t0
        import threading
        w = threading.Thread(target=watch)
        w.start()
        counts = dropin.hmmsearch_step(out, names, mol_type="prot", gencode=1, ingest=ing, devices=tuple(range(ngpu)),
                                       write_tbl=False, chunk_reads=a.chunk)
        t2 = time.perf_counter()
        ph = dict(dropin.hmmsearch_step.last_timing)
        stats = dropin.hmmsearch_step.last_stats
        nrec = ing.n
        w.join()
        ing.close()
        assert all(os.path.exists(os.path.join(out, "ok", f"extract_reads_{g}.ok")) for g in names)
        run = {"n_gpus": ngpu, "seconds_file_to_file": t2 - t0, "reads_per_s": nreads / (t2 - t0), "records_after_dedup": nrec,
               "hit_rows": int(sum(counts.values())), "phases": ph, "start_ingest_s": t1 - t0, "ingest_done_at_s": t_ing[0],
               "gpu_ms_total_last_chunk_per_ctx": [float(s["ms_total"]) for s in stats]}
        lim = max(("search_s", "finish_ingest_s", "report_s", "write_outputs_s", "read_hmms_s", "upload_profiles_s"), key=lambda k: ph.get(k, 0.0))
        run["limiter"] = lim
        res["runs"].append(run)
        print(json.dumps(run), file=sys.stderr, flush=True)
        if not a.keep:
            shutil.rmtree(out, ignore_errors=True)
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    if not a.keep:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
