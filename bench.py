#!/usr/bin/env python
"""bench.py -- PhyloAln search hot path on B200: reads/s and GCUPS (BASELINE.json metric).

Workload (BASELINE.json configs[1], "config2"): synthetic 150 bp Illumina-like reads, six-frame
translated, searched against 1,000 BUSCO-size protein profile HMMs.  One *step* = one resident
chunk of `--reads-per-step` reads against ALL profiles (the full 10 M-read config is 10e6 /
reads-per-step identical steps; throughput is per read, so it extrapolates linearly).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

value      reads/s with the read chunk already resident in HBM (translate + all filters + domain
           stage + host-side domain scoring), max over ranks, whole job.
e2e        same metric through the public C ABI with HOST buffers: H2D of the reads, translation,
           search, D2H of the hit records, every step.
roofline   dominant kernel = msv_pack_kernel (sticky-inf SSV + exact MSV re-run): algorithmic 4 integer lane-ops per DP
           cell x cells per launch / CUDA-event time, against the issue rate of the kernel's only per-cell
           instruction (HFMA2.RELU, two cells each), microbenchmarked in this run.
cpu_baseline  the CPU oracle (C restatement of the hmmsearch pipeline: striped AVX2 MSV and Viterbi filters as
           in HMMER, scalar Forward / domain definition; one OpenMP region over all host cores) on a bounded
           sample of the same workload with the same STATS lines; kind "port" (no hmmsearch binary exists in
           this image).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/s (six-frame 150bp reads vs 1000 protein HMMs, hmmsearch-equivalent pipeline)"
UNIT = "reads/s"


def build_profiles(n_hmm: int, seed: int):
    from phyloaln_b200 import synth
    rng = np.random.default_rng(seed)
    Ms = synth.busco_lengths(n_hmm, rng)
    return [synth.random_hmm(f"syn{i:04d}", int(Ms[i]), rng) for i in range(n_hmm)]


def make_reads(n: int, hmms, seed: int, frac: float = 0.01):
    from phyloaln_b200 import synth
    rng = np.random.default_rng(seed)
    reads = synth.random_reads(n, 150, rng)
    synth.plant_reads(reads, hmms, frac, rng)
    return reads


def calibrate_on_gpu(ctx, hmms, seed=42, n=200, abc="amino"):
    """STATS LOCAL for synthetic profiles, scored by the GPU kernels themselves (hmmbuild's job)."""
    import math
    from phyloaln_b200 import hmmbuild_lite as hb
    rng = np.random.default_rng(seed)
    s200 = hb.random_sequences(abc, n, 200, rng)
    s100 = hb.random_sequences(abc, n, 100, rng)
    ntot = ctx.load_proteins(s200 + s100, abc=0 if abc == "amino" else 1)
    for i, h in enumerate(hmms):
        r = ctx.score_all(i, ntot, want=("msv", "vit", "fwd"))
        lam = math.log(2.0) + 1.44 / (h.M * hb.mean_match_relent_bits(h))

        def loc(bits):
            return -math.log(float(np.mean(np.exp(-lam * np.asarray(bits, dtype=np.float64))))) / lam
        ln2 = math.log(2.0)
        mmu = loc((r["msv"][:n].astype(np.float64) - hb.null1(200)) / ln2)
        vmu = loc((r["vit"][:n].astype(np.float64) - hb.null1(200)) / ln2)
        gmu = loc((r["fwd"][n:].astype(np.float64) - hb.null1(100)) / ln2)
        tau = gmu - math.log(-math.log(1.0 - 0.04)) / lam
        h.stats = {"MSV": (round(mmu, 4), round(lam, 5)), "VITERBI": (round(vmu, 4), round(lam, 5)),
                   "FORWARD": (round(tau, 4), round(lam, 5))}
    for i, h in enumerate(hmms):
        ctx.set_evparam(i, h.evparams())
    ctx.commit()


class ClockSampler:
    def __init__(self, gpu_index: int):
        self.p = None
        self.path = f"/tmp/pa_clocks_{os.getpid()}.csv"
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            tok = [t.strip() for t in line.split(",")]
            if len(tok) < 7:
                continue
            try:
                sm.append(float(tok[0]))
                mx.append(float(tok[1]))
            except ValueError:
                continue
            for nm, v in zip(names, tok[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.remove(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        return out


CPU_PROFILE_STRIDE = 4   # the CPU arm runs every 4th profile of the set (same width distribution) and scales linearly


def cpu_reference_setup(hmms, threads: int):
    """Oracle profiles for the CPU arm, with THE SAME STATS lines as the GPU arm: profiles the GPU arm already calibrated keep
    their numbers; otherwise the oracle calibrates with the same seed, sample size and formulas (hmmbuild_lite.calibrate ==
    calibrate_on_gpu; GPU and oracle scores are bit-identical, so the rounded STATS come out equal)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    from phyloaln_b200 import hmmbuild_lite as hb
    sub = hmms[::CPU_PROFILE_STRIDE]

    def one(h):
        P = O.Profile.from_hmm(h)
        if not h.stats:
            hb.calibrate(h, lambda hm, seqs, P=P: P.score_all(seqs, simd=True), seed=42, n=200)
            P.set_stats(h.evparams())
        return P
    with ThreadPoolExecutor(max(1, threads)) as ex:      # ctypes releases the GIL inside the filters
        profs = list(ex.map(one, sub))
    return sub, profs


def cpu_reference_run(hmms, reads, threads: int, seconds_target: float = 15.0, setup=None):
    """The oracle pipeline on the host cores over a bounded sample: six-frame translation in C, then every sampled profile
    against every frame in one OpenMP region (orc_search_many: striped AVX2 MSV and Viterbi filters -- the way HMMER's own
    filters are vectorised -- scalar Forward / domain definition on the survivors).  Returns (reads/s vs ALL profiles,
    description, GCUPS)."""
    from oracle import oracle as O
    sub, profs = setup or cpu_reference_setup(hmms, threads)
    opts = O.default_opts(simd_msv=1, simd_vit=1)
    total_M, all_M = sum(h.M for h in sub), sum(h.M for h in hmms)
    # size the sample from a short probe
    probe = reads[:256]
    off = np.arange(len(probe) + 1, dtype=np.int64) * probe.shape[1]
    t0 = time.perf_counter()
    dsq, toff = O.translate_reads(probe.reshape(-1), off, 1, threads)
    O.search_many(profs, dsq, toff, opts, threads)
    dt = max(time.perf_counter() - t0, 1e-3)
    n_reads = int(min(len(reads), max(256, seconds_target * len(probe) / dt)))
    sample = reads[:n_reads]
    off = np.arange(n_reads + 1, dtype=np.int64) * sample.shape[1]
    t0 = time.perf_counter()
    dsq, toff = O.translate_reads(sample.reshape(-1), off, 1, threads)      # translation is part of the timed pipeline
    ndom, stages = O.search_many(profs, dsq, toff, opts, threads)
    dt = time.perf_counter() - t0
    cells = float(toff[-1]) * total_M
    gcups = cells / dt / 1e9
    reads_per_s = n_reads / (dt * all_M / total_M)   # scale the profile count linearly to the full set
    desc = (f"{n_reads} reads x {len(sub)} profiles (every {CPU_PROFILE_STRIDE}th of {len(hmms)}, scaled linearly), six-frame translation + "
            f"full pipeline in C, one OpenMP region over {threads} threads; MSV and Viterbi filters striped AVX2 "
            f"({'on' if O.lib().orc_have_avx2() else 'OFF - scalar'}), Forward/domain definition scalar; same STATS lines as the GPU arm; "
            f"{dt:.1f}s, {gcups / max(1, threads):.1f} GCUPS per thread, {int(ndom)} domains")
    cpu_reference_run.last_seconds = dt
    return reads_per_s, desc, gcups


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--hmms", type=int, default=1000)
    ap.add_argument("--reads-per-step", type=int, default=524288)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-region", action="store_true", help="cudaProfilerStart/Stop around the timed resident steps (ncu --profile-from-start off)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "config2: 150bp synthetic reads (1% planted homologs, 0.1% N) x BUSCO-size protein HMMs, six-frame",
              "n_hmm": args.hmms, "reads_per_step_per_gpu": args.reads_per_step, "read_len": 150,
              "frames": 6, "full_config_reads": 10_000_000,
              "l2": "inputs larger than L2: every step streams ~190 MB of per-profile score tables (1000 profiles) per L2-sized "
                    "super-tile of targets plus a fresh, different read chunk (164 MB of translated targets at 524,288 reads); "
                    "no explicit flush"}
    hmms = build_profiles(args.hmms, seed=2002)

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        reads = make_reads(max(2000, min(args.reads_per_step, 65536)), hmms, seed=1002)
        setup = cpu_reference_setup(hmms, threads)
        vals, gcs, secs = [], [], []
        desc = ""
        for s in range(args.warmup + args.steps):
            v, desc, gc = cpu_reference_run(hmms, reads, threads, seconds_target=max(4.0, 60.0 / max(1, args.warmup + args.steps)), setup=setup)
            if s >= args.warmup:
                vals.append(v)
                gcs.append(gc)
                secs.append(cpu_reference_run.last_seconds)
        v = float(np.mean(vals)) if vals else 0.0
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)) if secs else None,   # one step = the bounded sample in cpu_baseline.sample
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/s16/f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                                 "gcups": float(np.mean(gcs)) if gcs else None,
                                 "gcups_per_core": float(np.mean(gcs)) / threads if gcs else None,
                                 "what": "oracle port (builder's CPU restatement of the hmmsearch pipeline, SIMD filters), not HMMER itself"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    from phyloaln_b200 import capi
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = capi.Context(local_rank)
    for h in hmms:
        ctx.add_hmm(h, require_stats=False)
    ctx.commit()
    calibrate_on_gpu(ctx, hmms)

    # measured integer-pipe peak (thread-instructions/s); lane-ops = 2 per packed instruction
    dpx = {m: ctx.microbench_dpx(i) for i, m in enumerate(["viaddmnmx_s16x2_relu", "vimnmx3_s16x2", "ssv_mix_dpx", "viaddmnmx_s32",
                                                           "hfma2_relu", "ssv_mix_hfma2"])}
    # ceiling of the SSV inner loop (registers only, no table fetch): one half-rate HFMA2.RELU per two cells
    peak_cells = dpx["hfma2_relu"] * 1e9 * 2.0
    peak_cells_mix = dpx["ssv_mix_hfma2"] * 1e9 * 16.0 / 12.0   # the earlier loop: 8 HFMA2 + 4 VIMNMX3 per 16 cells
    peak_cells_dpx = dpx["ssv_mix_dpx"] * 1e9 * 16.0 / 12.0

    nsteps = args.warmup + args.steps
    chunks = [make_reads(args.reads_per_step, hmms, seed=1002 + 7919 * (rank * nsteps + s)) for s in range(nsteps)]
    offs = np.arange(args.reads_per_step + 1, dtype=np.uint64) * 150
    sumM = sum(h.M for h in hmms)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---------------- device-resident pass: translate kernel time + search -------------
    sampler = ClockSampler(local_rank)
    times, stats_acc = [], []
    launches_timed = 0
    for s in range(nsteps):
        l0 = ctx.stats()["kernel_launches"]
        if args.profile_region and s == args.warmup:
            capi.profiler_range(True)
        ctx.load_reads(chunks[s].reshape(-1), offs)            # reads resident after this call
        tr_ms = ctx.stats()["ms_translate"]
        barrier()
        t0 = time.perf_counter()
        hits, _, _ = ctx.search()
        dt = time.perf_counter() - t0 + tr_ms * 1e-3            # search (synchronous) + translation kernel
        barrier()
        if s >= args.warmup:
            times.append(dt)
            stats_acc.append(ctx.stats())
            launches_timed += stats_acc[-1]["kernel_launches"] - l0
    if args.profile_region:
        capi.profiler_range(False)
    # ---------------- end-to-end pass: host buffers in, hit records out ----------------
    e2e_times, d2h = [], 0
    for s in range(nsteps):
        barrier()
        t0 = time.perf_counter()
        ctx.load_reads(chunks[s].reshape(-1), offs)
        hits, model, aligned = ctx.search()
        dt = time.perf_counter() - t0
        barrier()
        if s >= args.warmup:
            e2e_times.append(dt)
            d2h = int(hits.nbytes + sum(map(len, model)) + sum(map(len, aligned)))
    clocks = sampler.stop()
    # ---------------- HBM-bound kernel: translation of a large resident chunk ----------------
    tr_roof = None
    if rank == 0:
        nbig = 1 << 20
        big = np.ascontiguousarray(np.tile(chunks[0], (nbig // args.reads_per_step + 1, 1))[:nbig])
        boffs = np.arange(nbig + 1, dtype=np.uint64) * 150
        best = 1e9
        for _ in range(3):
            ctx.load_reads(big.reshape(-1), boffs)
            best = min(best, ctx.stats()["ms_translate"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = 446.0 * nbig / (best * 1e-3) / 1e9
        tr_roof = {"bound": "hbm", "kernel": "translate_kernel", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                   "frac": ach / hbm_peak, "traffic": None, "reads": nbig, "ms": best,
                   "algorithmic_bytes_per_read": 446,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
        del big
    launches_total = launches_timed

    tot = float(np.sum(times))
    tot_e2e = float(np.sum(e2e_times))
    if dist is not None:
        import torch
        t = torch.tensor([tot, tot_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot, tot_e2e = float(t[0]), float(t[1])
    reads_total = args.reads_per_step * args.steps * world
    value = reads_total / tot
    e2e_value = reads_total / tot_e2e
    ms_msv = float(np.mean([s["ms_msv"] for s in stats_acc]))
    ms_stage = {k: float(np.mean([s[k] for s in stats_acc])) for k in
                ("ms_translate", "ms_msv", "ms_bias", "ms_vit", "ms_fwd", "ms_domain", "ms_domain_kernel", "ms_total")}
    cells_per_step = 296.0 * args.reads_per_step * sumM        # residues/read (50+49+49)*2 x sum of M
    padded_cells_per_step = 296.0 * args.reads_per_step * sum(16 * (h.M // 16 + 1) if h.M < 1024 else 64 * (h.M // 64 + 1) for h in hmms)
    msv_cells_s = cells_per_step / (ms_msv * 1e-3)
    gcups = cells_per_step * world / (tot / args.steps) / 1e9
    last = stats_acc[-1]
    sm_clk = (clocks["sm_mhz"] or 1965.0) * 1e6
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 MSV semantics evaluated in fp16x2 / s16 Viterbi: packed s16x2 upper bound + exact s32 DPX / f32", "data": "synthetic", "config": config,
            "gcups": gcups,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(chunks[0].nbytes + offs.nbytes),
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_total),
            "clocks": clocks,
            # The dominant kernel is bound by instruction issue on the FMA + ALU pipes, not by HBM or the tensor cores
            # (max-plus recurrence with loop-carried dependences; emission table resident in TMEM).  achieved / peak are
            # in the SURVEY 8d unit: algorithmic 4 integer lane-ops per DP cell.
            "roofline": {"bound": "int_dpx", "kernel": "msv_pack_kernel<16,2> + msv_tmem_kernel<17..32,2> (sticky-inf SSV: one fp16x2 HFMA2.RELU per two cells, no running maximum; packed profile tables in TMEM; inline exact MSV)",
                         "achieved": 4.0 * msv_cells_s / 1e12, "peak": 4.0 * peak_cells / 1e12,
                         "unit": "T 16-bit lane-op/s (algorithmic 4 per DP cell)",
                         "frac": msv_cells_s / peak_cells,
                         # dram__bytes_read.sum + dram__bytes_write.sum of msv_pack_kernel from ncu captures of this command
                         # (profiles/r1d_ncu_full_kernels.md): 131,072 reads/step 128.0 MB + 20.8 MB (--set full); 524,288 reads/step
                         # 1,426.3 MB + 83.9 MB with the L2 super-tile work order (metrics-only pass; 87.9 GB before it);
                         # 32,768 reads/step 38.6 MB (profiles/r2_ncu_msv_pack.md, --set full)
                         "traffic": {32768: 38.6e6, 131072: 148.8e6, 524288: 1510.2e6}.get(args.reads_per_step), "traffic_unit": "bytes per msv_pack_kernel launch",
                         "msv_gcups": msv_cells_s / 1e9,
                         "useful_cells_per_clk_per_sm": msv_cells_s / sm_clk / 148.0,
                         "padded_cells_per_clk_per_sm": padded_cells_per_step / (ms_msv * 1e-3) / sm_clk / 148.0,
                         "peak_source": "measured in this run (pa_microbench_dpx mode 4): issue rate of HFMA2.RELU (half rate on the FMA pipe), "
                                        "the only per-cell instruction of the SSV loop, x 2 cells, registers only",
                         "frac_of_hfma2_vimnmx3_mix_peak": msv_cells_s / peak_cells_mix,
                         "frac_of_dpx_only_mix_peak": msv_cells_s / peak_cells_dpx,
                         "tmem_read_B_per_clk_per_sm": 2.0 * padded_cells_per_step / (ms_msv * 1e-3) / sm_clk / 148.0,
                         "binding_resource": "the half-rate HFMA2 pipe (16 HFMA2.RELU per 1,024-cell warp row = 32 pipe cycles) and warp "
                                             "instruction issue (profiles/r2_ncu_msv_pack.md: fp16 pipe 86.7 %, issue active 76 %)",
                         "traffic_note": "DRAM traffic is not the bound (1.5 GB per 2.19 s launch at 524,288 reads/step, 149 MB per 0.55 s at 131,072): "
                                         "targets (one L2-sized super-tile at a time) + tables in, candidate records out; emission tables live in TMEM",
                         "ginstr_per_s": dpx},
            "roofline_translate": tr_roof,
            "stage_ms": ms_stage,
            "funnel": dict({k: int(last[k]) for k in ("n_pairs", "n_past_ssv", "n_past_msv", "n_past_bias", "n_past_vit", "n_past_fwd", "n_domains")},
                           n_vit_exact=int(last["vit_exact_pairs"]))}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                v, desc, gc = cpu_reference_run(hmms, chunks[0], os.cpu_count() or 1, seconds_target=12.0)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc,
                                        "gcups": gc, "gcups_per_core": gc / (os.cpu_count() or 1),
                                        "what": "oracle port (builder's CPU restatement of the hmmsearch pipeline, SIMD filters), not HMMER itself"}
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
