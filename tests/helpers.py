"""Shared builders for the test-suite (synthetic calibrated profiles, planted targets)."""
import numpy as np

from oracle import oracle as O
from phyloaln_b200 import hmmbuild_lite as hb
from phyloaln_b200 import synth


def calibrated_hmm(name, M, seed, bg_mix=0.25, abc="amino", n=60):
    """Random profile calibrated with the oracle (small n keeps the CPU suite fast)."""
    rng = np.random.default_rng(seed)
    h = synth.random_hmm(name, M, rng, abc=abc)
    if bg_mix != 0.5:  # sharper emissions -> planted fragments are found
        p = np.exp(-h.mat[1:])
        bg = hb.background(abc)
        d = (p - 0.5 * bg) / 0.5
        p2 = (1 - bg_mix) * d + bg_mix * bg
        p2 = np.clip(p2, 1e-6, None)
        p2 /= p2.sum(1, keepdims=True)
        h = hb.finish_hmm(name, abc, np.vstack([np.zeros((1, h.K)), p2]), np.exp(-h.ins), np.exp(-h.t))
    P = O.Profile.from_hmm(h)
    hb.calibrate(h, lambda hm, seqs: P.score_all(seqs), seed=seed, n=n)
    return h


def protein_targets(hmms, n, rng, Lmin=30, Lmax=70, plant_every=7):
    seqs = []
    for i in range(n):
        L = int(rng.integers(Lmin, Lmax + 1))
        s = hb.random_sequences("amino", 1, L, rng)[0]
        if plant_every and i % plant_every == 0:
            h = hmms[int(rng.integers(len(hmms)))]
            flen = min(h.M, max(8, L - 10))
            s0 = int(rng.integers(1, h.M - flen + 2))
            frag = synth.sample_protein(h, rng, s0, s0 + flen - 1)
            a = int(rng.integers(0, L - flen + 1)) if L > flen else 0
            s = (s[:a] + frag + s[a + flen:])[:max(L, flen)]
        seqs.append(s)
    return seqs
