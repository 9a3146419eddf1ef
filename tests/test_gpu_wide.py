"""GPU parity for profiles wider than 2,047 columns (csrc/wide.cuh): nucleotide / codon-mode references of ordinary
genes reach that width (the reference's own fixture tests/ref/OG0003977.fa is 2,049-2,052 nt) and hmmsearch
(/root/reference/lib/aln.py:664-667) has no limit.  Model-tiled MSV, multi-warp Viterbi, run-time-B Forward and the
domain stage must reproduce the oracle exactly as the register-resident kernels do for narrow profiles."""
import numpy as np
import pytest

from tests import helpers
from tests.test_gpu_search import compare_hits

pytestmark = pytest.mark.gpu


def _targets(abc, hmms, n, rng, Lmin, Lmax, plant_every=3):
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    seqs = []
    for i in range(n):
        L = int(rng.integers(Lmin, Lmax + 1))
        s = hb.random_sequences(abc, 1, L, rng)[0]
        if plant_every and i % plant_every == 0 and L > 12:
            h = hmms[int(rng.integers(len(hmms)))]
            flen = min(h.M, L - 5)
            a = int(rng.integers(1, h.M - flen + 2))
            frag = synth.sample_protein(h, rng, a, a + flen - 1)
            pos = int(rng.integers(0, L - len(frag) + 1))
            s = s[:pos] + frag + s[pos + len(frag):]
        seqs.append(s)
    return seqs


@pytest.mark.parametrize("abc,M", [("dna", 2048), ("dna", 2052), ("amino", 2111), ("dna", 3000), ("amino", 4096), ("dna", 4097),
                                   ("dna", 6000), ("amino", 8192)])
def test_raw_filter_scores_bit_exact_wide(oracle, gpu_ctx_factory, abc, M):
    """pa_score_all on one wide profile: exact MSV bytes (msv_exact_wide_kernel), exact Viterbi words (vit_wide_kernel,
    4 or 8 warps per pair) and the Forward score (fwd_wide_kernel) for every target, empty and 1-residue ones included."""
    rng = np.random.default_rng(M)
    h = helpers.calibrated_hmm(f"w{M}", M, seed=M, abc=abc, n=16, bg_mix=0.15)
    seqs = _targets(abc, [h], 90, rng, 1, 330, plant_every=2)
    seqs[5] = ""
    seqs[6] = seqs[6][:1]
    P = oracle.Profile.from_hmm(h)
    ctx = gpu_ctx_factory()
    idx = ctx.add_hmm(h)
    ctx.commit()
    n = ctx.load_proteins(seqs, abc=0 if abc == "amino" else 1)
    r = ctx.score_all(idx, n, want=("msv", "vit", "fwd"))
    for t, s in enumerate(seqs):
        if len(s) == 0:
            continue
        osc, oraw = P.msv(s)
        assert r["msv_raw"][t] == oraw, ("msv", t, len(s))
        assert r["msv"][t] == np.float32(osc)
        vsc, vraw = P.vit(s)
        assert r["vit_raw"][t] == vraw, ("vit", t, len(s))
        assert r["vit"][t] == np.float32(vsc)
        fsc = P.fwd(s)[0]
        assert r["fwd"][t] == np.float32(fsc), ("fwd", t, r["fwd"][t], fsc)
    ctx.close()


def _funnel(oracle, hmms, seqs, opts=None):
    want = np.zeros(6, dtype=np.int64)
    for h in hmms:
        _, tr = oracle.Profile.from_hmm(h).search(seqs, opts=opts, nthreads=8)
        want += np.bincount(tr["stage"], minlength=6)
    return want[1:].sum(), want[2:].sum(), want[3:].sum(), want[4:].sum()


@pytest.mark.parametrize("fullt", [False, True])
def test_search_wide_dna_profiles_nucleotide_reads(oracle, gpu_ctx_factory, fullt, monkeypatch):
    """Config 3 as SURVEY 8(d) defines it: DNA profiles of M = 3 x the protein widths (up to 6,000) next to narrow ones,
    read-sized nucleotide targets + reverse complements.  Hits, strings, scores and every funnel count equal the oracle's,
    so the tiled SSV pass flags a superset of the exact candidates and the 16-bit Viterbi bound drops no passing pair."""
    if fullt:   # the all-TMEM variant of the MSV rows (Q up to 26 words, 1,664-cell packs, all-TMEM tiles): off by default, kept exact
        monkeypatch.setenv("PA_MSV_FULLT", "1")
    rng = np.random.default_rng(2048)
    Ms = [300, 1100, 1500, 1663, 1664, 2047, 2048, 2500, 3500, 4100, 6000] if fullt else [300, 2047, 2048, 2500, 3500, 4100, 6000]
    hmms = [helpers.calibrated_hmm(f"wd{M}", M, seed=5000 + M, abc="dna", n=16, bg_mix=0.1) for M in Ms]
    reads = _targets("dna", hmms, 700, rng, 40, 260, plant_every=3)
    for i in range(0, len(reads), 6):
        reads[i] = oracle.revcomp(reads[i])
    reads[11] = reads[11][:10] + "NRY" + reads[11][13:]
    targets = [x for r in reads for x in (r, oracle.revcomp(r))]
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    assert ctx.load_read_list(reads, moltype=1) == len(targets)
    hits, model, aligned = ctx.search()
    n = compare_hits(oracle, hmms, targets, hits, model, aligned, pp=ctx.hit_pp(hits))
    assert n > 100
    assert len({int(x) for x in hits["hmm"]}) == len(Ms)
    st = ctx.stats()
    assert (st["n_past_msv"], st["n_past_bias"], st["n_past_vit"], st["n_past_fwd"]) == _funnel(oracle, hmms, targets)
    assert st["n_past_ssv"] >= st["n_past_msv"]
    ctx.close()


@pytest.mark.parametrize("F1,F2", [(0.02, 1e-3), (1.0, 1.0), (0.3, 1e-9)])
def test_search_wide_protein_profiles_long_targets(oracle, gpu_ctx_factory, F1, F2):
    """Wide protein profiles against transcript-like targets longer than one MSV tile (the +inf flag has to travel through
    several tile boundaries and rows), with default, wide-open and very strict filter thresholds."""
    rng = np.random.default_rng(77)
    Ms = [150, 2100, 2600, 4300]
    hmms = [helpers.calibrated_hmm(f"wp{M}", M, seed=7000 + M, n=16, bg_mix=0.15) for M in Ms]
    seqs = _targets("amino", hmms, 70, rng, 20, 2600, plant_every=2)
    opts = oracle.default_opts(F1=F1, F2=F2)
    ctx = gpu_ctx_factory()
    ctx.set_opts(F1=F1, F2=F2)
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    assert compare_hits(oracle, hmms, seqs, hits, model, aligned, opts=opts, pp=ctx.hit_pp(hits)) > 20
    st = ctx.stats()
    assert (st["n_past_msv"], st["n_past_bias"], st["n_past_vit"], st["n_past_fwd"]) == _funnel(oracle, hmms, seqs, opts)
    ctx.close()


def test_too_wide_profile_is_rejected(gpu_ctx_factory):
    from phyloaln_b200 import synth
    from phyloaln_b200.capi import PaError
    ctx = gpu_ctx_factory()
    h = synth.random_hmm("huge", 8193, np.random.default_rng(1), abc="dna")
    with pytest.raises(PaError, match="1..8192"):
        ctx.add_hmm(h, require_stats=False)
    ctx.close()


def test_documented_capacity_limits_are_loud_errors(gpu_ctx_factory):
    """65,535 profiles per context (Pass records carry a 16-bit profile index) and 65,535 distinct target lengths per chunk
    (16-bit length-table index): one more is an error with a message, -2 (bad argument) and -3 (split the chunk) respectively."""
    import ctypes as C
    from phyloaln_b200 import capi, synth
    from phyloaln_b200.capi import PaError
    ctx = gpu_ctx_factory()
    tiny = capi.prepare_hmm(synth.random_hmm("t", 1, np.random.default_rng(2)), require_stats=False)
    for _ in range(65535):
        ctx.add_hmm(tiny)
    with pytest.raises(PaError, match=r"\(-2\).*too many profiles"):
        ctx.add_hmm(tiny)
    ctx.close()
    ctx = gpu_ctx_factory()
    lens = np.arange(65536, dtype=np.uint64)                      # 65,536 distinct lengths, 2.1 GB of residues
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    aa = np.full(int(off[-1]), ord("A"), dtype=np.uint8)
    rc = ctx.L.pa_load_proteins(ctx.h, aa.ctypes.data, off.ctypes.data, len(lens), 0)
    assert rc == -3 and b"distinct target lengths" in ctx.L.pa_last_error(ctx.h)
    rc = ctx.L.pa_load_proteins(ctx.h, aa.ctypes.data, off.ctypes.data, len(lens) - 1, 0)   # 65,535 distinct lengths fit
    assert rc == len(lens) - 1
    ctx.close()
