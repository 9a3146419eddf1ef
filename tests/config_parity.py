"""Per-config parity on the real workloads (test infrastructure: the oracle is the checker, never the thing measured).

For BASELINE.json configs 2-5 the exact generator of tools/run_config.py (phyloaln_b200.synth.config_profiles /
config_chunk, same seeds) produces the full profile set and the first chunk; a subsample of that chunk and of the profile
set -- the narrowest and widest profiles, the ones that fill a 1,024-cell MSV pack almost alone, and the profiles with the
most hits -- is searched on the GPU and by the oracle WITH THE SAME STATS LINES (calibrated on the GPU as the benchmark
does) and compared: every filter-funnel count, the hit set, coordinates, alignment and PP strings bit-exact, scores within
0.01 bit, ln P within 1e-3 (the tolerances of BASELINE.json's north_star).  The same reads are also searched against the
FULL profile set (the benchmark's packs and work order) and the hits of the selected profiles must be identical.

  python -m tests.config_parity --config 3 [--units 4096] [--profiles 32] [--out profiles/r2_config_parity.json]
  python -m tests.config_parity --msv-impl [--reads 524288]     # TMEM path vs PA_MSV_IMPL=smem at bench size (config 2)

Reference call site of the stage being checked: /root/reference/lib/aln.py:664-667 (hmmsearch) and :482-570 (extract_reads).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

INT_FIELDS = ["ndom_in_seq", "dom_idx", "ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto", "multidomain_region"]
FLOAT_FIELDS = ["envsc", "domcorrection", "oasc", "dombias", "dom_bits", "seq_bits", "pre_bits", "seqbias", "fwdsc", "nullsc"]
BIT_TOL = 0.01       # bits (north_star)
LNP_TOL = 1e-3       # relative (north_star)


def _targets_of(O, nt, off, abc):
    """The target sequences the reference would search: six frames (lib/aln.py:249-258) or read + reverse complement (:241-244)."""
    raw = nt.tobytes().decode()
    reads = [raw[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    if abc == "dna":
        return [x for r in reads for x in (r, O.revcomp(r))]
    return [p for r in reads for _, p in O.six_frames(r, 1)]


def compare_with_oracle(O, hmms, targets, hits, model, aligned, pp, nthreads):
    """Raises AssertionError on the first difference; returns the summary of what was compared."""
    got = {(int(hits["hmm"][i]), int(hits["target"][i]), int(hits["dom_idx"][i])): i for i in range(len(hits))}
    nwant, max_bits, max_lnp = 0, 0.0, 0.0
    funnel = np.zeros(6, dtype=np.int64)
    for hi, h in enumerate(hmms):
        ohits, tr = O.Profile.from_hmm(h).search(targets, nthreads=nthreads)
        funnel += np.bincount(tr["stage"], minlength=6)
        for oh in ohits:
            nwant += 1
            key = (hi, oh["target"], oh["dom_idx"])
            assert key in got, f"hit missing on the GPU: profile {h.name} (M={h.M}) target {oh['target']} domain {oh['dom_idx']}"
            g = got[key]
            for f in INT_FIELDS:
                assert int(hits[f][g]) == oh[f], (h.name, key, f, int(hits[f][g]), oh[f])
            assert model[g] == oh["model"] and aligned[g] == oh["aligned"] and pp[g] == oh["pp"], (h.name, key, "alignment strings")
            for f in FLOAT_FIELDS:
                d = abs(float(hits[f][g]) - oh[f])
                max_bits = max(max_bits, d if f.endswith("bits") else d / 0.6931)   # the other fields are in nats
                assert d <= BIT_TOL * 0.6931 + 1e-5 * abs(oh[f]), (h.name, key, f, float(hits[f][g]), oh[f])
            for f in ("dom_lnP", "seq_lnP"):
                d = abs(float(hits[f][g]) - oh[f]) / max(1.0, abs(oh[f]))
                max_lnp = max(max_lnp, d)
                assert d <= LNP_TOL, (h.name, key, f)
    assert nwant == len(hits), f"the GPU reports {len(hits)} domains, the oracle {nwant}"
    return {"domains": nwant, "max_score_diff_bits": max_bits, "max_lnP_rel_diff": max_lnp,
            "funnel_oracle": {"n_past_msv": int(funnel[1:].sum()), "n_past_bias": int(funnel[2:].sum()),
                              "n_past_vit": int(funnel[3:].sum()), "n_past_fwd": int(funnel[4:].sum())}}


def select_profiles(hmms, hit_counts, n_sel):
    """Indices: the 4 narrowest, the 4 widest, the 2 nearest below the 1,024-cell pack limit, then the ones with most hits."""
    order = sorted(range(len(hmms)), key=lambda i: hmms[i].M)
    sel = list(order[:4]) + list(order[-4:])
    near = sorted((i for i in range(len(hmms)) if hmms[i].M < 1024), key=lambda i: -hmms[i].M)[:2]
    sel += [i for i in near if i not in sel]
    for i in np.argsort(-hit_counts, kind="stable"):
        if len(sel) >= n_sel:
            break
        if int(i) not in sel:
            sel.append(int(i))
    return sorted(sel[:n_sel])


def verify_config(config: int, n_units: int | None = None, n_sel: int = 32, device: int = 0, nthreads: int | None = None,
                  chunk_units: int | None = None, log=lambda *a: None):
    import bench
    from oracle import oracle as O
    from phyloaln_b200 import capi, synth
    nthreads = nthreads or os.cpu_count() or 1
    shape = synth.CONFIG_SHAPES[config]
    abc = shape["abc"]
    n_units = n_units or (4096 if shape["unit"] == "reads" else 384)
    t0 = time.time()
    rng = synth.config_rng(config)
    hmms = synth.config_profiles(config, rng)
    ctxA = capi.Context(device)
    for h in hmms:
        ctxA.add_hmm(h, require_stats=False)
    ctxA.commit()
    bench.calibrate_on_gpu(ctxA, hmms, abc=abc)                      # the STATS lines both sides use
    chunk_units = chunk_units or min(shape["chunk"], 32 * n_units)
    nt, off, planted = synth.config_chunk(config, chunk_units, rng, hmms)   # the first chunk of the real run
    moltype = 1 if abc == "dna" else 0
    ctxA.load_reads(nt, off, moltype=moltype)
    nf = ctxA.num_frames()
    hitsA, _, _ = ctxA.search()
    log(f"config {config}: {len(hmms)} profiles (M {min(h.M for h in hmms)}..{max(h.M for h in hmms)}), first chunk of {chunk_units} "
        f"{shape['unit']}: {len(hitsA)} domains on the GPU, {time.time() - t0:.1f}s")
    counts = np.bincount(hitsA["hmm"], minlength=len(hmms))
    sel = select_profiles(hmms, counts, n_sel)
    selset = set(sel)
    # units: half of the sample are those with a hit to a selected profile, the rest the head of the chunk (unbiased)
    hit_units = sorted({int(t) // nf for t, h in zip(hitsA["target"], hitsA["hmm"]) if int(h) in selset})[:n_units // 2]
    units = list(hit_units)
    seen = set(units)
    for u in range(chunk_units):
        if len(units) >= n_units:
            break
        if u not in seen:
            units.append(u)
    units.sort()
    raw = nt.tobytes()
    parts = [raw[int(off[u]):int(off[u + 1])] for u in units]
    sub_off = np.zeros(len(parts) + 1, dtype=np.uint64)
    sub_off[1:] = np.cumsum([len(p) for p in parts])
    sub_nt = np.frombuffer(b"".join(parts), dtype=np.uint8)
    sel_hmms = [hmms[i] for i in sel]
    # GPU, selected profiles only
    ctxB = capi.Context(device)
    for h in sel_hmms:
        ctxB.add_hmm(h)
    ctxB.commit()
    ntg = ctxB.load_reads(sub_nt, sub_off, moltype=moltype)
    hitsB, modelB, alignedB = ctxB.search()
    ppB = ctxB.hit_pp(hitsB)
    stB = ctxB.stats()
    ctxB.close()
    targets = _targets_of(O, sub_nt, sub_off, abc)
    assert ntg == len(targets)
    t1 = time.time()
    rep = compare_with_oracle(O, sel_hmms, targets, hitsB, modelB, alignedB, ppB, nthreads)
    funnel_gpu = {k: int(stB[k]) for k in ("n_past_msv", "n_past_bias", "n_past_vit", "n_past_fwd")}
    assert funnel_gpu == rep["funnel_oracle"], (funnel_gpu, rep["funnel_oracle"])
    log(f"  oracle on {len(units)} {shape['unit']} x {len(sel)} profiles: {rep['domains']} domains identical, funnel {funnel_gpu}, "
        f"{time.time() - t1:.1f}s")
    # GPU, full profile set (the benchmark's packs, tiles and work order) on the same reads
    ctxA.load_reads(sub_nt, sub_off, moltype=moltype)
    hitsA2, modelA2, alignedA2 = ctxA.search()
    ppA2 = ctxA.hit_pp(hitsA2)
    ctxA.close()
    remap = {g: k for k, g in enumerate(sel)}
    keep = [i for i in range(len(hitsA2)) if int(hitsA2["hmm"][i]) in remap]
    a = hitsA2[keep].copy()
    a["hmm"] = [remap[int(x)] for x in a["hmm"]]
    a["ali_off"] = 0
    b = hitsB.copy()
    b["ali_off"] = 0
    ka = np.lexsort((a["dom_idx"], a["target"], a["hmm"]))
    kb = np.lexsort((b["dom_idx"], b["target"], b["hmm"]))
    assert len(a) == len(b), ("full profile set vs selected profiles", len(a), len(b))
    for f in a.dtype.names:
        assert np.array_equal(a[f][ka], b[f][kb], equal_nan=(a[f].dtype.kind == "f")), ("full profile set vs selected profiles", f)
    for x, y in zip(ka, kb):
        i = keep[int(x)]
        assert modelA2[i] == modelB[int(y)] and alignedA2[i] == alignedB[int(y)] and ppA2[i] == ppB[int(y)]
    rep.update({"config": config, "workload": shape["what"], "n_profiles_total": len(hmms), "profiles_compared": len(sel),
                "profile_widths_compared": [h.M for h in sel_hmms], "units_compared": len(units), "unit": shape["unit"],
                "targets_compared": len(targets), "pairs_compared": len(targets) * len(sel), "funnel_gpu": funnel_gpu,
                "n_past_ssv_gpu": int(stB["n_past_ssv"]), "full_profile_set_hits_identical": True,
                "strings_identical": True, "tolerances": {"score_bits": BIT_TOL, "lnP_rel": LNP_TOL}, "ok": True,
                "seconds": time.time() - t0})
    return rep


def verify_msv_impl(n_reads: int = 524288, device: int = 0, log=lambda *a: None):
    """One bench-size step of config 2 through the TMEM kernels (packs, super-tile order) and through the shared-memory DPX
    kernel (PA_MSV_IMPL=smem): identical funnel and identical hit records."""
    import bench
    from phyloaln_b200 import capi, synth
    rng = synth.config_rng(2)
    hmms = synth.config_profiles(2, rng)
    nt, off, _ = synth.config_chunk(2, n_reads, rng, hmms)
    out = {}
    for impl in ("tmem", "smem"):
        if impl == "smem":
            os.environ["PA_MSV_IMPL"] = "smem"
        else:
            os.environ.pop("PA_MSV_IMPL", None)
        try:
            ctx = capi.Context(device)
        finally:
            os.environ.pop("PA_MSV_IMPL", None)
        for h in hmms:
            ctx.add_hmm(h, require_stats=not (impl == "tmem"))
        ctx.commit()
        if impl == "tmem":
            bench.calibrate_on_gpu(ctx, hmms)
        ctx.load_reads(nt, off)
        hits, model, aligned = ctx.search()
        st = ctx.stats()
        ctx.close()
        k = np.lexsort((hits["dom_idx"], hits["target"], hits["hmm"]))
        out[impl] = (hits[k], [model[i] for i in k], [aligned[i] for i in k],
                     {f: int(st[f]) for f in ("n_pairs", "n_past_msv", "n_past_bias", "n_past_vit", "n_past_fwd", "n_domains")}, float(st["ms_msv"]))
        log(f"  {impl}: funnel {out[impl][3]}, msv stage {out[impl][4]:.0f} ms")
    assert out["tmem"][3] == out["smem"][3], (out["tmem"][3], out["smem"][3])
    ha, hb_ = out["tmem"][0], out["smem"][0]
    assert len(ha) == len(hb_)
    for f in ha.dtype.names:
        if f == "ali_off":   # position in the string pool: depends on the order the survivors were appended
            continue
        bad = np.flatnonzero(~((ha[f] == hb_[f]) | ((ha[f] != ha[f]) & (hb_[f] != hb_[f]))))
        assert len(bad) == 0, (f, len(bad), ha[bad[:3]], hb_[bad[:3]])
    assert out["tmem"][1] == out["smem"][1] and out["tmem"][2] == out["smem"][2]
    return {"check": "config 2, one bench-size step: TMEM path (packs, super-tiles) vs shared-memory DPX kernel", "reads": n_reads,
            "n_profiles": len(hmms), "funnel": out["tmem"][3], "identical_pass1_sets_and_hits": True,
            "ms_msv_tmem": out["tmem"][4], "ms_msv_smem": out["smem"][4], "ok": True}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, nargs="*", default=[])
    ap.add_argument("--units", type=int, default=0)
    ap.add_argument("--profiles", type=int, default=32)
    ap.add_argument("--msv-impl", action="store_true")
    ap.add_argument("--reads", type=int, default=524288)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = []

    def log(*x):
        print(*x, file=sys.stderr, flush=True)
    for c in a.config:
        res.append(verify_config(c, n_units=a.units or None, n_sel=a.profiles, log=log))
    if a.msv_impl:
        res.append(verify_msv_impl(a.reads, log=log))
    print(json.dumps(res))
    if a.out:
        prev = []
        if os.path.exists(a.out):
            try:
                prev = json.load(open(a.out))
            except Exception:
                prev = []
        json.dump(prev + res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
