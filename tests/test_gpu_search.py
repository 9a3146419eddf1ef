"""GPU parity: the whole search (filters + domain definition + alignment) against the oracle."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

INT_FIELDS = ["ndom_in_seq", "dom_idx", "ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto", "multidomain_region"]
FLOAT_FIELDS = ["envsc", "domcorrection", "oasc", "dombias", "dom_bits", "seq_bits", "pre_bits", "seqbias", "fwdsc", "nullsc"]


def compare_hits(oracle, hmms, seqs, hits, model, aligned, opts=None, pp=None):
    """Every oracle hit must be reproduced exactly: set, coordinates, strings; scores within 0.01 bit."""
    got = {}
    for i in range(len(hits)):
        got[(int(hits["hmm"][i]), int(hits["target"][i]), int(hits["dom_idx"][i]))] = i
    nwant = 0
    for hi, h in enumerate(hmms):
        P = oracle.Profile.from_hmm(h)
        ohits, _ = P.search(seqs, opts=opts, nthreads=8, want_trace=False)
        for oh in ohits:
            nwant += 1
            key = (hi, oh["target"], oh["dom_idx"])
            assert key in got, f"missing hit {key}"
            g = got[key]
            for f in INT_FIELDS:
                assert int(hits[f][g]) == oh[f], (key, f, int(hits[f][g]), oh[f])
            assert model[g] == oh["model"], key
            assert aligned[g] == oh["aligned"], key
            if pp is not None:      # posterior-probability annotation of the alignment (HMMER's PP line)
                assert pp[g] == oh["pp"], key
            for f in FLOAT_FIELDS:
                assert abs(float(hits[f][g]) - oh[f]) <= 0.01 * 0.6931 + 1e-5 * abs(oh[f]), (key, f, float(hits[f][g]), oh[f])
            for f in ("dom_lnP", "seq_lnP"):
                assert abs(float(hits[f][g]) - oh[f]) <= 1e-3 * max(1.0, abs(oh[f])), (key, f)
    assert nwant == len(hits), (nwant, len(hits))
    return nwant


def test_search_matches_oracle_proteins(oracle, gpu_ctx_factory):
    rng = np.random.default_rng(77)
    hmms = [helpers.calibrated_hmm(f"h{i}", M, seed=10 + i) for i, M in enumerate([40, 90, 130, 200, 310])]
    seqs = helpers.protein_targets(hmms, 3000, rng, Lmin=25, Lmax=60, plant_every=9)
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    n = compare_hits(oracle, hmms, seqs, hits, model, aligned, pp=ctx.hit_pp(hits))
    assert n > 50
    ctx.close()


def test_search_matches_oracle_long_targets(oracle, gpu_ctx_factory):
    """Transcript-like targets (config 4 shape): several domains per target, rescaling, M > 512."""
    rng = np.random.default_rng(5)
    hmms = [helpers.calibrated_hmm(f"L{i}", M, seed=40 + i, bg_mix=0.15) for i, M in enumerate([150, 330, 600])]
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    seqs = []
    for i in range(150):
        L = int(rng.integers(200, 900))
        s = hb.random_sequences("amino", 1, L, rng)[0]
        for _ in range(int(rng.integers(0, 3))):
            h = hmms[int(rng.integers(len(hmms)))]
            a = int(rng.integers(1, h.M // 2))
            b = int(rng.integers(a + 30, h.M + 1))
            frag = synth.sample_protein(h, rng, a, b)
            pos = int(rng.integers(0, max(1, L - len(frag))))
            s = s[:pos] + frag + s[pos + len(frag):]
        seqs.append(s)
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    n = compare_hits(oracle, hmms, seqs, hits, model, aligned, pp=ctx.hit_pp(hits))
    assert n > 40
    assert (hits["ndom_in_seq"] > 1).any()
    ctx.close()


def test_search_six_frame_reads(oracle, gpu_ctx_factory):
    """Reads -> six-frame translation on the GPU -> search; oracle gets the frames from its own translator."""
    from phyloaln_b200 import synth
    rng = np.random.default_rng(9)
    hmms = [helpers.calibrated_hmm(f"r{i}", M, seed=60 + i, bg_mix=0.1) for i, M in enumerate([80, 160, 260])]
    reads = synth.random_reads(2000, 150, rng)
    synth.plant_reads(reads, hmms, 0.1, rng, sub_rate=0.02)
    rl = [r.tobytes().decode() for r in reads]
    frames = []
    for r in rl:
        frames.extend(p for _, p in oracle.six_frames(r, 1))
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    assert ctx.load_read_list(rl, moltype=0, gencode=1) == len(frames)
    hits, model, aligned = ctx.search()
    n = compare_hits(oracle, hmms, frames, hits, model, aligned, pp=ctx.hit_pp(hits))
    assert n > 30
    ctx.close()


def test_max_mode_and_nobias(oracle, gpu_ctx_factory):
    rng = np.random.default_rng(3)
    hmms = [helpers.calibrated_hmm("m0", 70, seed=91)]
    seqs = helpers.protein_targets(hmms, 200, rng, Lmin=30, Lmax=50, plant_every=4)
    for kw, ok in (({"do_max": True}, {"do_max": 1}), ({"do_bias": False}, {"do_bias": 0}), ({"do_null2": False}, {"do_null2": 0})):
        ctx = gpu_ctx_factory()
        ctx.add_hmm(hmms[0])
        ctx.set_opts(**kw)
        ctx.commit()
        ctx.load_proteins(seqs)
        hits, model, aligned = ctx.search()
        compare_hits(oracle, hmms, seqs, hits, model, aligned, opts=oracle.default_opts(**ok))
        ctx.close()


def test_nucleotide_mode_dna_profiles(oracle, gpu_ctx_factory):
    """Config 3 shape: DNA-alphabet profiles against nucleotide targets (read + its reverse complement)."""
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    rng = np.random.default_rng(31)
    hmms = [helpers.calibrated_hmm(f"d{i}", M, seed=300 + i, bg_mix=0.1, abc="dna") for i, M in enumerate([90, 240, 700])]
    reads = []
    for i in range(600):
        L = int(rng.integers(60, 200))
        s = hb.random_sequences("dna", 1, L, rng)[0]
        if i % 4 == 0:
            h = hmms[int(rng.integers(len(hmms)))]
            flen = min(h.M, L - 5)
            a = int(rng.integers(1, h.M - flen + 2))
            frag = synth.sample_protein(h, rng, a, a + flen - 1)
            pos = int(rng.integers(0, L - flen + 1))
            s = s[:pos] + frag + s[pos + flen:]
            if i % 8 == 0:
                s = oracle.revcomp(s)
        if i % 50 == 0:
            s = s[:10] + "NRY" + s[13:]
        reads.append(s)
    targets = []
    for r in reads:
        targets += [r, oracle.revcomp(r)]
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    assert ctx.load_read_list(reads, moltype=1) == len(targets)
    assert ctx.get_targets(len(targets), sum(map(len, targets))) == [t.upper() for t in targets]
    hits, model, aligned = ctx.search()
    n = compare_hits(oracle, hmms, targets, hits, model, aligned, pp=ctx.hit_pp(hits))
    assert n > 40
    # forward-only variant (-n / no_reverse)
    assert ctx.load_read_list(reads, moltype=1, no_reverse=True) == len(reads)
    hits, model, aligned = ctx.search()
    assert compare_hits(oracle, hmms, reads, hits, model, aligned, pp=ctx.hit_pp(hits)) > 20
    ctx.close()


def test_codon_mode_first_frame_only(oracle, gpu_ctx_factory):
    """--codon: only frame 1 of the forward strand is translated (lib/aln.py:249-253, main.py:196-198)."""
    from phyloaln_b200 import synth
    rng = np.random.default_rng(12)
    hmms = [helpers.calibrated_hmm(f"c{i}", M, seed=70 + i, bg_mix=0.1) for i, M in enumerate([60, 120])]
    reads = synth.random_reads(1500, 150, rng)
    synth.plant_reads(reads, hmms, 0.2, rng, sub_rate=0.02)
    rl = [r.tobytes().decode() for r in reads]
    frames = [p for r in rl for _, p in oracle.six_frames(r, 11, no_reverse=True, codon=True)]
    assert len(frames) == len(rl)
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    assert ctx.load_read_list(rl, moltype=0, gencode=11, no_reverse=True, codon_only=True) == len(frames)
    hits, model, aligned = ctx.search()
    assert compare_hits(oracle, hmms, frames, hits, model, aligned, pp=ctx.hit_pp(hits)) > 10
    ctx.close()


def test_search_very_long_targets_and_profiles(oracle, gpu_ctx_factory):
    """Config 4 extremes: targets up to 1,700 aa, profiles beyond 1,024 columns (hybrid TMEM + smem MSV)."""
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    rng = np.random.default_rng(8)
    hmms = [helpers.calibrated_hmm(f"V{i}", M, seed=140 + i, bg_mix=0.15, n=30) for i, M in enumerate([420, 1100, 1500])]
    seqs = []
    for i in range(60):
        L = int(rng.integers(166, 1700))
        s = hb.random_sequences("amino", 1, L, rng)[0]
        if i % 2 == 0:
            h = hmms[int(rng.integers(len(hmms)))]
            a = int(rng.integers(1, h.M // 2))
            b = int(rng.integers(a + 40, min(h.M, a + 400) + 1))
            frag = synth.sample_protein(h, rng, a, b)
            if len(frag) < L:
                pos = int(rng.integers(0, L - len(frag)))
                s = s[:pos] + frag + s[pos + len(frag):]
        seqs.append(s)
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    assert compare_hits(oracle, hmms, seqs, hits, model, aligned, pp=ctx.hit_pp(hits)) > 15
    ctx.close()


def test_multidomain_regions_are_resolved_by_trace_clustering(oracle, gpu_ctx_factory):
    """Tandem copies of a profile fragment inside one region trip the rt3 test; the region is split into
    envelopes by 200 stochastic tracebacks + single-linkage clustering, identically on GPU and oracle."""
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    rng = np.random.default_rng(11)
    hmms = [helpers.calibrated_hmm("tandA", 120, seed=77, bg_mix=0.1, n=40), helpers.calibrated_hmm("tandB", 300, seed=78, bg_mix=0.1, n=40)]
    seqs = []
    for t in range(80):
        h = hmms[t % 2]
        a = int(rng.integers(1, h.M // 4))
        b = int(rng.integers(h.M // 2, h.M + 1))
        ncopy = 2 + (t % 5 == 0)
        parts = [hb.random_sequences("amino", 1, int(rng.integers(5, 40)), rng)[0]]
        for c in range(ncopy):
            parts.append(synth.sample_protein(h, rng, a, b))
            if t % 2:
                parts.append(hb.random_sequences("amino", 1, int(rng.integers(0, 12)), rng)[0])
        parts.append(hb.random_sequences("amino", 1, int(rng.integers(5, 40)), rng)[0])
        seqs.append("".join(parts))
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    n = compare_hits(oracle, hmms, seqs, hits, model, aligned, pp=ctx.hit_pp(hits))
    multi = hits[hits["multidomain_region"] == 1]
    assert len(multi) >= 40 and n >= 150
    # a flagged region really was split: several domains share it
    assert (multi["ndom_in_seq"] >= 2).all()
    ctx.close()


def test_packed_profiles_boundaries(oracle, gpu_ctx_factory):
    """Profile packing (msv_pack_kernel): widths around the 16-cell half-lane and the 1,024-cell pack limits,
    many tiny profiles in one pack, a profile that fills a pack alone, one that is too wide to be packed."""
    rng = np.random.default_rng(21)
    Ms = [1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 127, 128, 500, 1007, 1008, 1023, 1024, 1040]
    hmms = [helpers.calibrated_hmm(f"pk{M}", M, seed=600 + M, bg_mix=0.1, n=24) for M in Ms]
    seqs = helpers.protein_targets(hmms, 700, rng, Lmin=5, Lmax=70, plant_every=2)
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    n = compare_hits(oracle, hmms, seqs, hits, model, aligned, pp=ctx.hit_pp(hits))
    assert n > 100
    assert len(set(int(x) for x in hits["hmm"])) >= 12   # hits spread over the packed profiles
    # the funnel counts agree with the oracle as well (every pair, candidate or not)
    st = ctx.stats()
    past_msv = 0
    for h in hmms:
        _, tr = oracle.Profile.from_hmm(h).search(seqs, nthreads=8)
        past_msv += int((tr["stage"] >= 1).sum())
    assert st["n_past_msv"] == past_msv
    ctx.close()


@pytest.mark.parametrize("F1", [0.02, 0.3, 1.0])
def test_sticky_ssv_long_targets_in_packs_and_loose_thresholds(oracle, gpu_ctx_factory, F1):
    """The sticky-inf SSV pass (no running maximum): targets longer than the 1,024-cell packed model (the +inf flag wraps
    around the pack several times), many narrow members per pack, and MSV thresholds so loose that the candidate threshold
    Cp collapses to 1 or below (scale +inf / every pair a candidate).  Hits and every funnel count must equal the oracle's."""
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    rng = np.random.default_rng(31)
    Ms = [18, 25, 40, 40, 55, 70, 90, 120, 160, 200, 260, 310]
    hmms = [helpers.calibrated_hmm(f"st{i}", M, seed=900 + i, bg_mix=0.1, n=24) for i, M in enumerate(Ms)]
    seqs = []
    for i in range(90):
        L = int(rng.integers(30, 120)) if i % 3 else int(rng.integers(900, 2600))
        s = hb.random_sequences("amino", 1, L, rng)[0]
        if i % 2 == 0:
            h = hmms[int(rng.integers(len(hmms)))]
            frag = synth.sample_protein(h, rng, 1, h.M)
            if len(frag) < L:
                pos = int(rng.integers(0, L - len(frag)))
                s = s[:pos] + frag + s[pos + len(frag):]
        seqs.append(s)
    ctx = gpu_ctx_factory()
    ctx.set_opts(F1=F1)
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    opts = oracle.default_opts(F1=F1)
    assert compare_hits(oracle, hmms, seqs, hits, model, aligned, opts=opts, pp=ctx.hit_pp(hits)) >= 15
    st = ctx.stats()
    want = np.zeros(6, dtype=np.int64)
    for h in hmms:
        _, tr = oracle.Profile.from_hmm(h).search(seqs, opts=opts, nthreads=8)
        want += np.bincount(tr["stage"], minlength=6)
    assert (st["n_past_msv"], st["n_past_bias"], st["n_past_vit"], st["n_past_fwd"]) == (want[1:].sum(), want[2:].sum(), want[3:].sum(), want[4:].sum())
    assert st["n_past_ssv"] >= st["n_past_msv"]
    ctx.close()


@pytest.mark.parametrize("F2", [1.0, 0.2, 1e-9])
def test_viterbi_threshold_extremes(oracle, gpu_ctx_factory, F2):
    """The 16-bit Viterbi upper-bound pass only drops pairs whose bound fails the pair's own threshold: F2 = 1 (threshold
    -inf: every pair goes to the exact kernel), a loose and a very strict F2 must all give the oracle's cascade and hits."""
    rng = np.random.default_rng(41)
    hmms = [helpers.calibrated_hmm(f"v{i}", M, seed=300 + i, bg_mix=0.1, n=30) for i, M in enumerate([35, 64, 129, 300, 520])]
    seqs = helpers.protein_targets(hmms, 1500, rng, Lmin=20, Lmax=90, plant_every=5)
    ctx = gpu_ctx_factory()
    ctx.set_opts(F2=F2)
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    hits, model, aligned = ctx.search()
    opts = oracle.default_opts(F2=F2)
    assert compare_hits(oracle, hmms, seqs, hits, model, aligned, opts=opts) > 30
    st = ctx.stats()
    want = np.zeros(6, dtype=np.int64)
    for h in hmms:
        _, tr = oracle.Profile.from_hmm(h).search(seqs, opts=opts, nthreads=8)
        want += np.bincount(tr["stage"], minlength=6)
    assert (st["n_past_bias"], st["n_past_vit"], st["n_past_fwd"]) == (want[2:].sum(), want[3:].sum(), want[4:].sum())
    assert st["vit_exact_pairs"] <= st["n_past_bias"]
    ctx.close()
