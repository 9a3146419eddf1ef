"""Host-side profile construction ("hmmbuild-lite") and calibration.

The reference shells out to HMMER ``hmmbuild`` (/root/reference/lib/aln.py:56-91); the north star
keeps HMM building on the host.  ``hmmbuild`` is not available in this image, so test fixtures
and synthetic benchmark profiles are produced here: Henikoff
position-based weights, symfrac match-column assignment, entropy-weighted effective sequence
number, Dirichlet transition priors, background-Dirichlet emission prior (HMMER's 9-component
mixture is not reproduced), E-value calibration by simulation.  Files written by
``hmmio.write_hmm`` are valid HMMER3/f, so real ``hmmbuild`` output is interchangeable.
"""
from __future__ import annotations

import math

import numpy as np

from .hmmio import AMINO, DNA, HMM

AMINO_BG = np.array([0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198,
                     0.0590092, 0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639,
                     0.0540978, 0.0683364, 0.0540687, 0.0673417, 0.0114135, 0.0304133])
DNA_BG = np.full(4, 0.25)
_DEGEN = {"amino": {"B": "DN", "J": "IL", "Z": "EQ", "O": "K", "U": "C", "X": AMINO},
          "dna": {"R": "AG", "Y": "CT", "M": "AC", "K": "GT", "S": "CG", "W": "AT", "H": "ACT", "B": "CGT",
                  "V": "ACG", "D": "AGT", "N": DNA, "U": "T", "X": DNA}}
# Dirichlet transition priors
_PRI_M = np.array([0.7939, 0.0278, 0.0135])
_PRI_I = np.array([0.1551, 0.1331])
_PRI_D = np.array([0.9002, 0.5630])


def background(abc: str) -> np.ndarray:
    return AMINO_BG if abc == "amino" else DNA_BG


def _nlp(p):
    with np.errstate(divide="ignore"):
        return -np.log(np.asarray(p, dtype=np.float64))


def _resvec(ch: str, abc: str, alphabet: str, bg: np.ndarray):
    """Fractional count vector of one alignment character, or None for a gap."""
    ch = ch.upper()
    if ch in "-.~_*":
        return None
    v = np.zeros(len(alphabet))
    if ch in alphabet and not (abc == "dna" and ch in _DEGEN["dna"] and ch not in DNA):
        v[alphabet.index(ch)] = 1.0
        return v
    mem = _DEGEN[abc].get(ch, alphabet)
    idx = [alphabet.index(c) for c in mem]
    v[idx] = bg[idx] / bg[idx].sum()
    return v


def henikoff_weights(rows: list[str]) -> np.ndarray:
    n, alen = len(rows), len(rows[0])
    w = np.zeros(n)
    for c in range(alen):
        col = [r[c].upper() for r in rows]
        counts = {}
        for ch in col:
            if ch not in "-.~":
                counts[ch] = counts.get(ch, 0) + 1
        if not counts:
            continue
        r = len(counts)
        for s, ch in enumerate(col):
            if ch in counts:
                w[s] += 1.0 / (r * counts[ch])
    if w.sum() <= 0:
        return np.ones(n)
    return w * n / w.sum()


def _parameterize(cm, ct, scale, bg, prior_a):
    """Counts (scaled by `scale`) + priors -> probabilities. cm (M+1,K), ct (M+1,7)."""
    M = cm.shape[0] - 1
    mat = np.zeros_like(cm)
    mat[1:] = (cm[1:] * scale + prior_a * bg) / (cm[1:].sum(1, keepdims=True) * scale + prior_a)
    t = np.zeros_like(ct)
    c = ct * scale
    t[:, 0:3] = (c[:, 0:3] + _PRI_M) / (c[:, 0:3].sum(1, keepdims=True) + _PRI_M.sum())
    t[:, 3:5] = (c[:, 3:5] + _PRI_I) / (c[:, 3:5].sum(1, keepdims=True) + _PRI_I.sum())
    t[:, 5:7] = (c[:, 5:7] + _PRI_D) / (c[:, 5:7].sum(1, keepdims=True) + _PRI_D.sum())
    # node M: no delete/insert continuation; node 0: no delete state D_0
    t[M, 0:3] = [t[M, 0] / (t[M, 0] + t[M, 1]), t[M, 1] / (t[M, 0] + t[M, 1]), 0.0]
    t[M, 5:7] = [1.0, 0.0]
    t[0, 5:7] = [1.0, 0.0]
    return mat, t


def _mean_relent(mat, bg):
    p = mat[1:]
    with np.errstate(divide="ignore", invalid="ignore"):
        h = np.where(p > 0, p * np.log2(p / bg), 0.0).sum(1)
    return float(h.mean())


def occupancy(t: np.ndarray):
    """Match/insert occupancies from transition probabilities (p7_hmm_CalculateOccupancy)."""
    M = t.shape[0] - 1
    mocc = np.zeros(M + 1)
    iocc = np.zeros(M + 1)
    mocc[1] = t[0, 1] + t[0, 0]
    for k in range(2, M + 1):
        mocc[k] = mocc[k - 1] * (t[k - 1, 0] + t[k - 1, 1]) + (1.0 - mocc[k - 1]) * t[k - 1, 5]
    iocc[0] = t[0, 1] / max(t[0, 3], 1e-12)
    for k in range(1, M):
        iocc[k] = mocc[k] * t[k, 1] / max(t[k, 3], 1e-12)
    return mocc, iocc


def finish_hmm(name, abc, mat, ins, t, cons=None, mapcols=None, nseq=1, effn=1.0) -> HMM:
    """Assemble an HMM object (probabilities in, -ln p out) and set COMPO."""
    M = mat.shape[0] - 1
    mocc, iocc = occupancy(t)
    compo = (mocc[1:, None] * mat[1:]).sum(0) + (iocc[:, None] * ins).sum(0)
    compo = compo / compo.sum()
    matn = _nlp(mat)
    matn[0] = math.inf
    return HMM(name=name, abc=abc, mat=matn, ins=_nlp(ins), t=_nlp(t), compo=_nlp(compo), cons=cons,
               nseq=nseq, effn=effn, map=mapcols, rf=None if mapcols is None else "x" * M)


def build_from_alignment(name: str, seqs: dict[str, str], abc: str = "amino", symfrac: float = 0.5,
                         prior_a: float = 1.0):
    """Build an (uncalibrated) profile from an alignment. Returns (HMM, match_columns [0-based])."""
    alphabet = AMINO if abc == "amino" else DNA
    bg = background(abc)
    K = len(alphabet)
    rows = list(seqs.values())
    alen = len(rows[0])
    if any(len(r) != alen for r in rows):
        raise ValueError("alignment rows differ in length")
    n = len(rows)
    w = henikoff_weights(rows)
    vecs = [[_resvec(ch, abc, alphabet, bg) for ch in r] for r in rows]
    frac = np.array([sum(w[s] for s in range(n) if vecs[s][c] is not None) for c in range(alen)]) / w.sum()
    mcols = [c for c in range(alen) if frac[c] >= symfrac]
    M = len(mcols)
    if M == 0:
        raise ValueError("no match columns")
    col2node = {c: k + 1 for k, c in enumerate(mcols)}
    cm = np.zeros((M + 1, K))
    ct = np.zeros((M + 1, 7))
    for s in range(n):
        # state path: node index + state in {M, I, D}; begins at B (treated as M_0)
        prev_k, prev_st = 0, "M"
        for c in range(alen):
            v = vecs[s][c]
            if c in col2node:
                k = col2node[c]
                st = "M" if v is not None else "D"
                if st == "M":
                    cm[k] += w[s] * v
                ct[prev_k, {"MM": 0, "MD": 2, "IM": 3, "DM": 5, "DD": 6}[prev_st + st]] += w[s] if (prev_st + st) in (
                    "MM", "MD", "IM", "DM", "DD") else 0.0
                prev_k, prev_st = k, st
            elif v is not None:
                if prev_st == "D":
                    continue  # D->I not allowed in Plan 7; drop the residue (rare)
                ct[prev_k, 1 if prev_st == "M" else 4] += w[s]
                prev_st = "I"
    # entropy weighting: find scale on counts giving the target mean relative entropy
    ere = 0.59 if abc == "amino" else 0.62
    target = max(ere, (45.0 - math.log2(2.0 / (M * (M + 1.0)))) / M)
    lo, hi = 1e-4, 1.0
    mat, t = _parameterize(cm, ct, hi, bg, prior_a)
    if _mean_relent(mat, bg) > target:
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            mat, t = _parameterize(cm, ct, mid, bg, prior_a)
            if _mean_relent(mat, bg) > target:
                hi = mid
            else:
                lo = mid
        scale = 0.5 * (lo + hi)
    else:
        scale = 1.0
    mat, t = _parameterize(cm, ct, scale, bg, prior_a)
    ins = np.tile(bg, (M + 1, 1))
    thr = 0.5 if abc == "amino" else 0.9
    cons = "".join(alphabet[int(np.argmax(mat[k]))] if mat[k].max() >= thr else alphabet[int(np.argmax(mat[k]))].lower()
                   for k in range(1, M + 1))
    hmm = finish_hmm(name, abc, mat, ins, t, cons=cons, mapcols=[c + 1 for c in mcols], nseq=n, effn=scale * n)
    return hmm, mcols


def mean_match_relent_bits(hmm: HMM) -> float:
    bg = background(hmm.abc)
    return _mean_relent(np.vstack([np.zeros((1, hmm.K)), np.exp(-hmm.mat[1:])]), bg)


def null1(L: int) -> float:
    p1 = L / (L + 1.0)
    return L * math.log(p1) + math.log(1.0 - p1)


def random_sequences(abc: str, n: int, L: int, rng: np.random.Generator) -> list[str]:
    alphabet = AMINO if abc == "amino" else DNA
    bg = background(abc)
    idx = rng.choice(len(alphabet), size=(n, L), p=bg / bg.sum())
    arr = np.frombuffer(alphabet.encode(), dtype=np.uint8)[idx]
    return [row.tobytes().decode() for row in arr]


def calibrate(hmm: HMM, score_fn, seed: int = 42, n: int = 200) -> HMM:
    """Set STATS LOCAL MSV/VITERBI/FORWARD.

    score_fn(hmm, list_of_ascii_sequences) -> (msv_nats, vit_nats, fwd_nats) numpy arrays; the
    product passes its GPU scorer, fixture generation passes the CPU oracle.
    """
    rng = np.random.default_rng(seed)
    M = hmm.M
    lam = math.log(2.0) + 1.44 / (M * mean_match_relent_bits(hmm))

    def gumbel_loc(bits):
        return -math.log(float(np.mean(np.exp(-lam * np.asarray(bits, dtype=np.float64))))) / lam

    s200 = random_sequences(hmm.abc, n, 200, rng)
    msv, vit, _ = score_fn(hmm, s200)
    mmu = gumbel_loc((np.asarray(msv) - null1(200)) / math.log(2.0))
    vmu = gumbel_loc((np.asarray(vit) - null1(200)) / math.log(2.0))
    s100 = random_sequences(hmm.abc, n, 100, rng)
    _, _, fwd = score_fn(hmm, s100)
    gmu = gumbel_loc((np.asarray(fwd) - null1(100)) / math.log(2.0))
    tau = gmu - math.log(-math.log(1.0 - 0.04)) / lam
    hmm.stats = {"MSV": (round(mmu, 4), round(lam, 5)), "VITERBI": (round(vmu, 4), round(lam, 5)),
                 "FORWARD": (round(tau, 4), round(lam, 5))}
    return hmm
