"""GPU parity: translation and the integer/float filters against the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def test_translation_matches_oracle(oracle, gpu_ctx_factory):
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACGTACGTACGTACGTNRYKMSWBDHVacgtn", dtype=np.uint8)
    reads = []
    for i in range(400):
        L = int(rng.integers(0, 160)) if i % 10 else int(rng.integers(0, 5))
        reads.append(letters[rng.integers(0, len(letters), L)].tobytes().decode())
    reads[3] = "TAATAGTGATARTRATAN" + reads[3]
    # 'U': read as T on the forward strand; Bio.Seq.reverse_complement (>= 1.80) has no U in its DNA table and leaves it in
    # place, so it is read as T on the reverse strand as well (oracle: orc_revcomp) -- one pinned behaviour on both sides
    reads[4] = "AUGUUUGCUUAAUAGUGA" + reads[4]
    reads[5] = reads[5][:7] + "uUaUg" + reads[5][7:]
    assert oracle.revcomp("AUGC") == "GCUT" and [p for _, p in oracle.six_frames("AUGUUUGCU", 1)] == ["MFA", "CL", "VC", "CFL", "AF", "LS"]
    for table in (1, 5, 11):
        ctx = gpu_ctx_factory()
        n = ctx.load_read_list(reads, moltype=0, gencode=table)
        assert n == 6 * len(reads)
        got = ctx.get_targets(n, 2 * sum(map(len, reads)) + 64)
        k = 0
        for r in reads:
            for _, prot in oracle.six_frames(r, table):
                assert got[k] == prot, (table, r, k)
                k += 1
        ctx.close()
    # codon mode: frame 1 only, no reverse; dna mode: read + reverse complement
    ctx = gpu_ctx_factory()
    n = ctx.load_read_list(reads, moltype=0, gencode=1, no_reverse=True, codon_only=True)
    assert n == len(reads)
    got = ctx.get_targets(n, sum(map(len, reads)) + 64)
    for r, g in zip(reads, got):
        assert g == oracle.translate(r, 1)
    n = ctx.load_read_list(reads, moltype=1)
    got = ctx.get_targets(n, 2 * sum(map(len, reads)) + 64)
    for i, r in enumerate(reads):
        up = r.upper()
        assert got[2 * i] == up.replace("U", "T") and got[2 * i + 1] == oracle.revcomp(up).replace("U", "T")   # the DNA alphabet prints U as T
    ctx.close()


@pytest.mark.parametrize("longest", [150, 1024, 1025, 3000])
def test_translation_frame_sets_and_read_lengths(oracle, gpu_ctx_factory, longest):
    """Six-frame, forward-only and codon-only modes on ragged chunks whose longest read sits below, at and above the
    staging sizes of translate_kernel (shared-memory image for reads <= 1,024 nt, generic path beyond)."""
    rng = np.random.default_rng(longest)
    letters = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTNRYacgt", dtype=np.uint8)
    reads = [letters[rng.integers(0, len(letters), int(rng.integers(0, min(longest, 200) + 1)))].tobytes().decode() for _ in range(333)]
    reads[17] = letters[rng.integers(0, len(letters), longest)].tobytes().decode()
    reads[0], reads[1], reads[2] = "", "A", "AC"
    ctx = gpu_ctx_factory()
    for kw, frames in (({}, lambda r: [p for _, p in oracle.six_frames(r, 4)]),
                       ({"no_reverse": True}, lambda r: [p for _, p in oracle.six_frames(r, 4)][:3]),
                       ({"no_reverse": True, "codon_only": True}, lambda r: [oracle.translate(r, 4)])):
        n = ctx.load_read_list(reads, moltype=0, gencode=4, **kw)
        want = [p for r in reads for p in frames(r)]
        assert n == len(want)
        assert ctx.get_targets(n, 2 * sum(map(len, reads)) + 64) == want
    ctx.close()


def test_invalid_nucleotide_is_an_error(gpu_ctx_factory):
    from phyloaln_b200.capi import PaError
    ctx = gpu_ctx_factory()
    with pytest.raises(PaError):
        ctx.load_read_list(["ACGTACGTAC", "ACGTXACGT"], moltype=0)
    with pytest.raises(PaError):   # next to a read that takes the generic long-read path
        ctx.load_read_list(["ACGT" * 300, "ACGT!ACGT"], moltype=0)
    assert ctx.load_read_list(["ACGTACGTAC", "AC"], moltype=0) == 12
    ctx.close()


@pytest.mark.parametrize("M", [1, 3, 33, 64, 65, 130, 257, 450, 700, 1023, 1100, 2047])
def test_raw_filter_scores_bit_exact(oracle, gpu_ctx_factory, M):
    rng = np.random.default_rng(100 + M)
    h = helpers.calibrated_hmm(f"syn{M}", M, seed=M)
    seqs = helpers.protein_targets([h], 300, rng, Lmin=1, Lmax=90, plant_every=5)
    seqs[7] = "*" + seqs[7] + "*X"
    seqs[8] = "BJZOUX" + seqs[8]
    seqs[9] = ""
    P = oracle.Profile.from_hmm(h)
    ctx = gpu_ctx_factory()
    idx = ctx.add_hmm(h)
    ctx.commit()
    n = ctx.load_proteins(seqs)
    r = ctx.score_all(idx, n)
    for t, s in enumerate(seqs):
        if len(s) == 0:
            continue
        osc, oraw = P.msv(s)
        assert r["msv_raw"][t] == oraw, ("msv", t, s)
        assert r["msv"][t] == np.float32(osc)
        vsc, vraw = P.vit(s)
        assert r["vit_raw"][t] == vraw, ("vit", t, s)
        assert r["vit"][t] == np.float32(vsc)
        fsc = P.fwd(s)[0]
        assert r["fwd"][t] == np.float32(fsc), ("fwd", t, r["fwd"][t], fsc)
        assert abs(r["bias"][t] - P.bias(s)) <= 1e-5 * max(1.0, abs(P.bias(s)))
    ctx.close()


def test_filter_cascade_counts_match_oracle(oracle, gpu_ctx_factory):
    rng = np.random.default_rng(77)
    hmms = [helpers.calibrated_hmm(f"h{i}", M, seed=10 + i) for i, M in enumerate([40, 90, 130, 200, 310])]
    seqs = helpers.protein_targets(hmms, 3000, rng, Lmin=25, Lmax=60, plant_every=9)
    ctx = gpu_ctx_factory()
    for h in hmms:
        ctx.add_hmm(h)
    ctx.commit()
    ctx.load_proteins(seqs)
    ctx.search()
    st = ctx.stats()
    want = np.zeros(6, dtype=np.int64)
    for h in hmms:
        P = oracle.Profile.from_hmm(h)
        _, tr = P.search(seqs, nthreads=8)
        want += np.bincount(tr["stage"], minlength=6)
    past_msv = want[1:].sum()
    past_bias = want[2:].sum()
    past_vit = want[3:].sum()
    past_fwd = want[4:].sum()
    assert (st["n_past_msv"], st["n_past_bias"], st["n_past_vit"], st["n_past_fwd"]) == (past_msv, past_bias, past_vit, past_fwd)
    assert past_fwd > 50  # the planted fragments must exercise the whole cascade
    ctx.close()
