"""CPU tests of the oracle: known-answer vectors, brute-force enumeration, invariants and the
committed golden fixtures (what stands in for the missing HMMER goldens)."""
import gzip
import hashlib
import itertools
import math
import os

import numpy as np
import pytest

from tests import helpers

GOLD = os.path.join(os.path.dirname(__file__), "golden", "config1")
GROUPS = ["OG0003531", "OG0003709", "OG0003820", "OG0003977", "OG0004212"]
STD = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"


def test_translation_known_answers(oracle):
    # all 64 codons of the standard table, base order TCAG (NCBI gc.prt)
    for i, (a, b, c) in enumerate(itertools.product("TCAG", repeat=3)):
        assert oracle.translate(a + b + c, 1) == STD[i]
    # table differences typed from the NCBI listing
    assert oracle.translate("TGAATAAGAAGG", 2) == "WM**"      # vertebrate mito
    assert oracle.translate("TGAATAAGAAGG", 5) == "WMSS"      # invertebrate mito
    assert oracle.translate("TGACTG", 4) == "WL" and oracle.translate("AAAAGA", 9) == "NS"
    assert oracle.translate("TAATAGTGA", 11) == "***"
    # Biopython ambiguity rules
    assert oracle.translate("GCNCTNTCNCCNCGNACNGTNGGN", 1) == "ALSPRTVG"
    assert oracle.translate("RAYSARMTHTARTRATANNNN", 1) == "BZJ**XX"
    assert oracle.translate("acgtac", 1) == "TY" and oracle.translate("ACGTA", 1) == "T" and oracle.translate("AC", 1) == ""
    assert oracle.translate("AUGUUU", 1) == "MF"
    with pytest.raises(ValueError):
        oracle.translate("ACGTXA", 1)
    # six frames of the first DMELA read
    read = "TATAAATGCCATAAATATTAGATAAGTCGATAATTAACAAAAATTTTGTGGACTTTGCAAGCTTAGTTTAGCAGTATATTTCGTAGTAAATTATGTGAAC"
    want = ["YKCHKY*ISR*LTKILWTLQA*FSSIFRSKLCE", "INAINIR*VDN*QKFCGLCKLSLAVYFVVNYVN", "*MP*ILDKSIINKNFVDFASLV*QYIS**IM*",
            "VHIIYYEIYC*TKLAKSTKFLLIIDLSNIYGIY", "FT*FTTKYTAKLSLQSPQNFC*LSTYLIFMAFI", "SHNLLRNILLN*ACKVHKIFVNYRLI*YLWHL"]
    assert [p for _, p in oracle.six_frames(read, 1)] == want
    assert [s for s, _ in oracle.six_frames(read, 1)] == ["_pos1", "_pos2", "_pos3", "_rev_pos1", "_rev_pos2", "_rev_pos3"]


def test_reverse_complement(oracle):
    assert oracle.revcomp("ACGTRYMKSWHBVDNacgt") == "acgtNHBVDWSMKRYACGT"
    rng = np.random.default_rng(0)
    s = "".join("ACGTNRYKMSWBDHV"[i] for i in rng.integers(0, 15, 200))
    assert oracle.revcomp(oracle.revcomp(s)) == s


def _raw_reads():
    from phyloaln_b200 import dropin
    return dropin.read_raw_fasta(os.path.join(GOLD, "all.raw.fasta.gz"))


def test_target_fasta_matches_reference_raw2target(oracle):
    """all.target.fasta written by the reference's raw2target (dictionary semantics) vs the oracle frames."""
    ids, nt, off = _raw_reads()
    assert len(ids) == 2481 and len(nt) > 0          # 2,481 distinct ids
    raw = nt.tobytes().decode()
    out = []
    for i, rid in enumerate(ids):
        for suf, prot in oracle.six_frames(raw[int(off[i]):int(off[i + 1])], 1):
            out.append(f">{rid}{suf}\n{prot}\n")
    text = "".join(out)
    assert len(out) == 14886                           # Z of config 1
    assert text.startswith(open(os.path.join(GOLD, "target_head.fasta")).read())
    assert hashlib.sha256(text.encode()).hexdigest() == open(os.path.join(GOLD, "target.sha256")).read().strip()
    assert open(os.path.join(GOLD, "total_counts.tsv")).read() == "DMELA\t1340\nDWILL\t8916\n"


def test_hmm_file_readers_agree(oracle):
    from phyloaln_b200 import hmmio
    for g in GROUPS:
        path = os.path.join(GOLD, "ref_hmm", g + ".hmm")
        h = hmmio.read_hmm(path)
        P1, P2 = oracle.Profile.from_file(path), oracle.Profile.from_hmm(h)
        assert P1.M == P2.M == h.M and h.abc == "amino" and len(h.consensus()) == h.M
        L = oracle.lib()
        for k in (1, h.M // 2, h.M):
            for x in (0, 7, 19, 26, 27):
                assert L.orc_profile_msc(P1.p, x, k) == L.orc_profile_msc(P2.p, x, k)
                assert L.orc_profile_rbv(P1.p, x, k) == L.orc_profile_rbv(P2.p, x, k)
            for t in range(8):
                a, b = L.orc_profile_tsc(P1.p, k - 1, t), L.orc_profile_tsc(P2.p, k - 1, t)
                assert a == b or (math.isinf(a) and math.isinf(b))
        s = "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ"
        assert P1.msv(s) == P2.msv(s) and P1.vit(s) == P2.vit(s) and P1.fwd(s) == P2.fwd(s)


def _brute_force(P, L_, seq, viterbi):
    """Sum / max over ALL paths of the multihit local Plan-7 model for a tiny profile (nats)."""
    lib = P.L
    M, Lq = P.M, len(seq)
    d = P.digitize(seq)
    nj = 1.0
    pmove = (2.0 + nj) / (Lq + 2.0 + nj)
    tloop, tmove = math.log(1 - pmove), math.log(pmove)
    NEG = -math.inf

    def tsc(k, t):
        return lib.orc_profile_tsc(P.p, k, t)

    def msc(i, k):
        return lib.orc_profile_msc(P.p, int(d[i]), k)
    comb = max if viterbi else (lambda a, b: a if b == NEG else b if a == NEG else max(a, b) + math.log1p(math.exp(-abs(a - b))))
    from functools import lru_cache

    @lru_cache(None)
    def core(i, k, st):
        """score of completing the sequence from state st at (i residues consumed, node k) up to T"""
        if st == "M":
            r = comb(NEG, special(i, "E"))                                   # M_k -> E
            if k < M and i < Lq:
                r = comb(r, tsc(k, 0) + msc(i + 1, k + 1) + core(i + 1, k + 1, "M"))   # MM
                r = comb(r, tsc(k, 1) + core(i + 1, k, "I"))                          # MI (insert odds 1)
            if k < M:
                r = comb(r, tsc(k, 2) + core(i, k + 1, "D"))                          # MD
            return r
        if st == "I":
            r = NEG
            if i < Lq:
                r = comb(r, tsc(k, 3) + msc(i + 1, k + 1) + core(i + 1, k + 1, "M"))
                r = comb(r, tsc(k, 4) + core(i + 1, k, "I"))
            return r
        if st == "D":
            r = special(i, "E")                                              # D_k -> E
            if k < M and i < Lq:
                r = comb(r, tsc(k, 5) + msc(i + 1, k + 1) + core(i + 1, k + 1, "M"))
            if k < M:
                r = comb(r, tsc(k, 6) + core(i, k + 1, "D"))
            return r
        raise AssertionError

    @lru_cache(None)
    def special(i, st):
        if st == "E":
            return comb(math.log(0.5) + special(i, "C"), math.log(0.5) + special(i, "J"))
        if st == "C":
            r = tmove if i == Lq else NEG
            if i < Lq:
                r = comb(r, tloop + special(i + 1, "C"))
            return r
        if st == "J":
            r = tmove + special(i, "B")
            if i < Lq:
                r = comb(r, tloop + special(i + 1, "J"))
            return r
        if st == "N":
            r = tmove + special(i, "B")
            if i < Lq:
                r = comb(r, tloop + special(i + 1, "N"))
            return r
        if st == "B":
            r = NEG
            if i < Lq:
                for k in range(1, M + 1):
                    r = comb(r, tsc(k - 1, 7) + msc(i + 1, k) + core(i + 1, k, "M"))
            return r
        raise AssertionError
    return special(0, "N")


@pytest.mark.parametrize("M,L", [(1, 1), (2, 3), (3, 4), (4, 6)])
def test_forward_viterbi_against_path_enumeration(oracle, M, L):
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    rng = np.random.default_rng(17 * M + L)
    h = synth.random_hmm("tiny", M, rng)
    P = oracle.Profile.from_hmm(h)
    for seq in hb.random_sequences("amino", 4, L, rng):
        assert abs(P.generic_forward(seq) - _brute_force(P, None, seq, False)) < 1e-9
        assert abs(P.generic_viterbi(seq) - _brute_force(P, None, seq, True)) < 1e-9
        assert abs(P.fwd(seq)[0] - P.generic_forward(seq)) < 2e-5
        assert abs(P.bwd(seq) - P.generic_forward(seq)) < 2e-5


def test_filter_invariants(oracle):
    from phyloaln_b200 import hmmbuild_lite as hb
    rng = np.random.default_rng(4)
    h = helpers.calibrated_hmm("inv", 150, seed=4)
    P = oracle.Profile.from_hmm(h)
    seqs = helpers.protein_targets([h], 60, rng, Lmin=20, Lmax=400, plant_every=3)
    for s in seqs:
        f, b, g = P.fwd(s)[0], P.bwd(s), P.generic_forward(s)
        assert abs(f - g) < 1e-4 * max(1, abs(g)) and abs(b - g) < 1e-4 * max(1, abs(g))   # Forward == Backward
        # the 16-bit Viterbi filter tracks the exact Viterbi score within its documented -3 nat / rounding slack
        v, gv = P.vit(s)[0], P.generic_viterbi(s)
        assert (v > 0 and gv > 20.0) if math.isinf(v) else (gv - 3.2 <= v <= gv + 0.1)
        assert gv <= g + 1e-6                                                               # Viterbi <= Forward
        bt, et, mo = P.domain_decoding(s)
        assert abs(bt[-1] - et[-1]) < 1e-3 and (mo > -1e-4).all() and (mo < 1 + 1e-4).all()
    # a strong homolog needs rescaling (score >> 13 bits) and must stay finite and consistent
    big = "".join(helpers.synth.consensus_protein(h)) * 3
    f, e, m = P.fwd(big)
    assert e > 0 and abs(f - P.generic_forward(big)) < 1e-3 * abs(f) and abs(P.bwd(big) - f) < 1e-3 * abs(f)
    assert P.msv(big)[1] == -1 and math.isinf(P.msv(big)[0])                                # MSV overflow -> +inf
    assert P.vit("*" * 10)[0] < -40.0   # -32768 + entry score: saturating words are not a sticky -inf (A.5)


def test_statistics_functions(oracle):
    L = oracle.lib()
    assert abs(L.orc_gumbel_surv(-9.0, -10.0, 0.7) - (1 - math.exp(-math.exp(-0.7)))) < 1e-12
    assert L.orc_exp_surv(1.0, 2.0, 0.7) == 1.0 and abs(L.orc_exp_surv(5.0, 2.0, 0.7) - math.exp(-2.1)) < 1e-15
    assert L.orc_exp_logsurv(1.0, 2.0, 0.7) == 0.0 and abs(L.orc_exp_logsurv(5.0, 2.0, 0.7) + 2.1) < 1e-12
    assert abs(L.orc_null1(100) - (100 * math.log(100 / 101) + math.log(1 / 101))) < 1e-4


def test_oracle_reproduces_config1_golden(oracle):
    """Regression pin: oracle pipeline + report on config 1 == committed hit_info.tsv (all 777 rows)."""
    from oracle import report as orep
    ids, nt, off = _raw_reads()
    raw = nt.tobytes().decode()
    names, seqs = [], []
    for i, rid in enumerate(ids):
        for suf, prot in oracle.six_frames(raw[int(off[i]):int(off[i + 1])], 1):
            names.append(rid + suf)
            seqs.append(prot)
    total = 0
    for g in GROUPS:
        P = oracle.Profile.from_file(os.path.join(GOLD, "ref_hmm", g + ".hmm"))
        hits, _ = P.search(seqs, nthreads=os.cpu_count() or 1, want_trace=False)
        rows = orep.rows_from_hits(g, orep.apply_report(hits, names, len(names)), names)
        text = "".join("\t".join(str(x) for x in r) + "\n" for r in rows)
        assert text == open(os.path.join(GOLD, "hit_info", g + ".hit_info.tsv")).read(), g
        total += len(rows)
    assert total == 777


def test_alignment_strings_are_consistent(oracle):
    """Trace validity: coordinates, gap bookkeeping and residues of every reported alignment."""
    h = helpers.calibrated_hmm("ali", 120, seed=8, bg_mix=0.1)
    rng = np.random.default_rng(8)
    seqs = helpers.protein_targets([h], 400, rng, Lmin=30, Lmax=80, plant_every=3)
    # indels inside planted fragments
    seqs = [s[:20] + "WW" + s[20:] if i % 6 == 0 else (s[:25] + s[28:] if i % 6 == 3 else s) for i, s in enumerate(seqs)]
    P = oracle.Profile.from_hmm(h)
    hits, _ = P.search(seqs, nthreads=4, want_trace=False)
    assert len(hits) > 40
    cons = h.consensus()
    nins = ndel = 0
    for d in hits:
        m, a, s = d["model"], d["aligned"], seqs[d["target"]]
        assert len(m) == len(a) and m[0] != "." and a[0] != "-" and m[-1] != "." and a[-1] != "-"
        assert sum(c != "." for c in m) == d["hmmto"] - d["hmmfrom"] + 1
        assert a.replace("-", "").upper() == s[d["iali"] - 1:d["jali"]]
        assert "".join(c for c in m if c != ".").upper() == cons[d["hmmfrom"] - 1:d["hmmto"]].upper()
        assert d["ienv"] <= d["iali"] <= d["jali"] <= d["jenv"]
        nins += "." in m
        ndel += "-" in a
    assert nins > 0 and ndel > 0


def test_simd_msv_equals_scalar_msv(oracle):
    """The AVX2 striped MSV (bench.py's CPU baseline) returns the scalar filter's byte and score."""
    import numpy as np
    from tests import helpers
    rng = np.random.default_rng(4)
    for M in (1, 31, 32, 33, 200, 470, 2047):
        h = helpers.calibrated_hmm(f"s{M}", M, seed=900 + M, n=8)
        P = oracle.Profile.from_hmm(h)
        seqs = helpers.protein_targets([h], 120, rng, Lmin=1, Lmax=120, plant_every=3)
        seqs.append("*" + seqs[0] + "X*")
        for s in seqs:
            got = P.msv_simd(s)
            if got is None:
                pytest.skip("no AVX2 on this host")
            assert got == P.msv(s), (M, s)
    # and through the pipeline: same hits with and without the SIMD filter
    a, _ = P.search(seqs, opts=oracle.default_opts(simd_msv=1), nthreads=2)
    b, _ = P.search(seqs, nthreads=2)
    assert [(x["target"], x["iali"], x["jali"], x["seq_bits"]) for x in a] == [(x["target"], x["iali"], x["jali"], x["seq_bits"]) for x in b]


def test_simd_viterbi_equals_scalar_viterbi(oracle):
    """The AVX2 striped Viterbi filter (lazy D->D sweeps) returns the scalar filter's xC word and score, for amino and DNA
    profiles, widths around the 16-lane stripe boundaries, planted fragments (long D chains across lanes) and overflow."""
    import numpy as np
    from phyloaln_b200 import hmmbuild_lite as hb, synth
    from tests import helpers
    rng = np.random.default_rng(6)
    for abc, M in (("amino", 1), ("amino", 15), ("amino", 16), ("amino", 17), ("amino", 200), ("amino", 470), ("dna", 333), ("amino", 2100), ("dna", 800)):
        h = helpers.calibrated_hmm(f"v{M}", M, seed=700 + M, n=8, abc=abc, bg_mix=0.1)
        P = oracle.Profile.from_hmm(h)
        seqs = [hb.random_sequences(abc, 1, int(rng.integers(1, 160)), rng)[0] for _ in range(60)]
        for _ in range(25):   # fragments with deletions in the middle: D->D runs that cross stripe lanes
            a = int(rng.integers(1, max(2, M // 2)))
            b = int(rng.integers(a, M + 1))
            frag = synth.sample_protein(h, rng, a, b)
            if len(frag) > 12:
                cut = int(rng.integers(2, len(frag) - 2))
                frag = frag[:cut] + frag[cut + int(rng.integers(1, max(2, min(40, len(frag) - cut - 1)))):]
            seqs.append(frag)
        if M >= 200:
            seqs.append(synth.sample_protein(h, rng, 1, M) * 3)   # overflow of the 16-bit range
        for s in seqs:
            got = P.vit_simd(s)
            if got is None:
                pytest.skip("no AVX2 on this host")
            assert got == P.vit(s), (abc, M, len(s))


def test_search_many_and_translate_reads_equal_the_per_profile_calls(oracle):
    """bench.py's CPU baseline driver (one OpenMP region over all profiles, SIMD filters, C six-frame translation) counts
    the same funnel and domains as per-profile scalar searches over Python-translated frames."""
    import numpy as np
    from phyloaln_b200 import synth
    from tests import helpers
    rng = np.random.default_rng(12)
    hmms = [helpers.calibrated_hmm(f"m{i}", M, seed=50 + i, bg_mix=0.1, n=24) for i, M in enumerate([60, 140, 333])]
    reads = synth.random_reads(500, 150, rng)
    synth.plant_reads(reads, hmms, 0.2, rng, sub_rate=0.02)
    reads[7, 10] = ord("N")
    reads[8, 11] = ord("r")
    off = np.arange(len(reads) + 1, dtype=np.int64) * 150
    dsq, toff = oracle.translate_reads(reads.reshape(-1), off, 1, nthreads=3)
    frames = [p for r in reads for _, p in oracle.six_frames(r.tobytes().decode(), 1)]
    assert len(toff) == len(frames) + 1 and [int(b - a) for a, b in zip(toff[:-1], toff[1:])] == [len(f) for f in frames]
    cat = "".join(frames).encode()
    want = np.zeros(len(cat), dtype=np.uint8)
    oracle.lib().orc_digitize(0, cat, len(cat), want.ctypes.data)
    assert np.array_equal(dsq[:len(cat)], want)
    profs = [oracle.Profile.from_hmm(h) for h in hmms]
    nd, st = oracle.search_many(profs, dsq, toff, oracle.default_opts(simd_msv=1, simd_vit=1), nthreads=4)
    ndom, stages = 0, np.zeros(6, dtype=np.int64)
    for P in profs:
        hits, tr = P.search(frames, nthreads=4)
        ndom += len(hits)
        stages += np.bincount(tr["stage"], minlength=6)
    assert nd == ndom and st.tolist() == stages.tolist() and nd > 30
