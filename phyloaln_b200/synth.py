"""Seeded synthetic workloads: profiles, reads and transcripts.

Used by bench.py and the tests; nothing here is on the search path itself.
"""
from __future__ import annotations

import numpy as np

from .hmmbuild_lite import AMINO_BG, DNA_BG, finish_hmm
from .hmmio import AMINO, DNA, HMM

# one codon per amino acid for back-translation (standard code)
_CODON = {"A": "GCT", "C": "TGT", "D": "GAT", "E": "GAA", "F": "TTT", "G": "GGT", "H": "CAT", "I": "ATT",
          "K": "AAA", "L": "CTG", "M": "ATG", "N": "AAT", "P": "CCT", "Q": "CAA", "R": "CGT", "S": "TCT",
          "T": "ACT", "V": "GTT", "W": "TGG", "Y": "TAT"}


def random_hmm(name: str, M: int, rng: np.random.Generator, abc: str = "amino") -> HMM:
    """BUSCO-like random profile: Dirichlet match emissions mixed 50/50 with background."""
    bg = AMINO_BG if abc == "amino" else DNA_BG
    K = len(bg)
    mat = np.zeros((M + 1, K))
    d = rng.dirichlet(0.3 * bg * K, size=M)
    mat[1:] = 0.5 * d + 0.5 * bg
    mat[1:] /= mat[1:].sum(1, keepdims=True)
    t = np.zeros((M + 1, 7))
    mi = rng.uniform(0.005, 0.025, M + 1)
    md = rng.uniform(0.005, 0.025, M + 1)
    t[:, 0], t[:, 1], t[:, 2] = 1.0 - mi - md, mi, md
    ii = rng.uniform(0.3, 0.5, M + 1)
    t[:, 3], t[:, 4] = 1.0 - ii, ii
    dd = rng.uniform(0.25, 0.45, M + 1)
    t[:, 5], t[:, 6] = 1.0 - dd, dd
    t[M, 0:3] = [1.0 - mi[M], mi[M], 0.0]
    t[M, 5:7] = [1.0, 0.0]
    t[0, 5:7] = [1.0, 0.0]
    ins = np.tile(bg, (M + 1, 1))
    return finish_hmm(name, abc, mat, ins, t)


def busco_lengths(n: int, rng: np.random.Generator, mu: float = 420.0, sigma: float = 0.5, lo: int = 100,
                  hi: int = 2000) -> np.ndarray:
    return np.clip(np.rint(rng.lognormal(np.log(mu), sigma, n)), lo, hi).astype(int)


def sample_protein(hmm: HMM, rng: np.random.Generator, start: int = 1, end: int | None = None) -> str:
    """Emit match states start..end of the profile (no indels) -> a homologous fragment."""
    end = end or hmm.M
    p = np.exp(-hmm.mat[start:end + 1])
    p /= p.sum(1, keepdims=True)
    alphabet = AMINO if hmm.abc == "amino" else DNA
    return "".join(alphabet[rng.choice(len(alphabet), p=row)] for row in p)


def consensus_protein(hmm: HMM, start: int = 1, end: int | None = None) -> str:
    """Most probable residue of match states start..end."""
    end = end or hmm.M
    alphabet = AMINO if hmm.abc == "amino" else DNA
    return "".join(alphabet[i] for i in np.argmin(hmm.mat[start:end + 1], axis=1))


def back_translate(prot: str) -> str:
    return "".join(_CODON[a] for a in prot)


def random_reads(n: int, length: int, rng: np.random.Generator, n_frac: float = 0.001) -> np.ndarray:
    """(n, length) uint8 ASCII matrix of iid ACGT with a sprinkle of N."""
    arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, length))]
    if n_frac > 0:
        mask = rng.random((n, length)) < n_frac
        arr = np.where(mask, np.uint8(ord("N")), arr)
    return np.ascontiguousarray(arr)


def plant_reads(reads: np.ndarray, hmms: list[HMM], frac: float, rng: np.random.Generator, sub_rate: float = 0.05):
    """Overwrite a fraction of reads with back-translated profile-consensus fragments."""
    n, length = reads.shape
    idx = rng.choice(n, size=int(n * frac), replace=False)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for r in idx:
        h = hmms[rng.integers(len(hmms))]
        naa = length // 3
        if h.M < naa + 2:
            continue
        s = int(rng.integers(1, h.M - naa + 1))
        nt = back_translate(consensus_protein(h, s, s + naa - 1))
        off = int(rng.integers(0, length - len(nt) + 1))
        b = bytearray(reads[r].tobytes())
        b[off:off + len(nt)] = nt.encode()
        for i in np.nonzero(rng.random(length) < sub_rate)[0]:
            b[i] = b"ACGT"[rng.integers(4)]
        if rng.random() < 0.5:
            b = bytearray(bytes(b).translate(comp)[::-1])
        reads[r] = np.frombuffer(bytes(b), dtype=np.uint8)
    return idx


# ---------------------------------------------------------------------------------------------
# The BASELINE.json workloads, shared by tools/run_config.py, the per-config
# parity checks (tests/config_parity.py) and the file-to-file benchmark (tools/dropin_bench.py)
# ---------------------------------------------------------------------------------------------
CONFIG_SHAPES = {
    2: dict(units=10_000_000, n_hmm=1000, abc="amino", unit="reads", chunk=131072,
            what="10 M x 150 bp reads, six-frame, vs 1,000 BUSCO-size protein HMMs"),
    3: dict(units=10_000_000, n_hmm=1000, abc="dna", unit="reads", chunk=131072,
            what="10 M x 150 bp reads (+ reverse complement) vs 1,000 DNA HMMs, M = 3 x the protein widths (300..6,000)"),
    4: dict(units=200_000, n_hmm=3000, abc="amino", unit="transcripts", chunk=65536,
            what="200 k transcripts of 0.5-5 kb (30 % carry one or two planted domains), six-frame, vs 3,000 protein HMMs"),
    5: dict(units=100_000_000, n_hmm=3000, abc="amino", unit="reads", chunk=131072,
            what="100 M x 150 bp reads, six-frame, vs 3,000 protein HMMs"),
}


def config_rng(config: int) -> np.random.Generator:
    return np.random.default_rng(3000 + config)


def config_profiles(config: int, rng: np.random.Generator, n_hmm: int | None = None) -> list[HMM]:
    """The profile set of one BASELINE.json config (uncalibrated; consumes `rng` exactly as tools/run_config.py does)."""
    shape = CONFIG_SHAPES[config]
    n_hmm = n_hmm or shape["n_hmm"]
    Ms = busco_lengths(n_hmm, rng)
    if config == 3:
        Ms = 3 * Ms   # nucleotide models of the same genes: 300..6,000 columns
    return [random_hmm(f"c{config}_{i:04d}", int(Ms[i]), rng, abc=shape["abc"]) for i in range(n_hmm)]


def config_chunk(config: int, n: int, rng: np.random.Generator, hmms: list[HMM]):
    """One chunk of n reads / transcripts of the config: (nt uint8, off uint64[n+1], number planted)."""
    abc = CONFIG_SHAPES[config]["abc"]
    n_hmm = len(hmms)
    planted = 0
    if config in (2, 3, 5):
        reads = random_reads(n, 150, rng)
        if abc == "amino":
            planted = len(plant_reads(reads, hmms, 0.01, rng))
        else:  # fragments sampled from the DNA profiles themselves
            for r in rng.choice(n, size=n // 100, replace=False):
                h = hmms[int(rng.integers(n_hmm))]
                s = int(rng.integers(1, h.M - 150 + 1)) if h.M > 151 else 1
                frag = sample_protein(h, rng, s, min(h.M, s + 149)).encode()
                reads[r, :len(frag)] = np.frombuffer(frag, dtype=np.uint8)
                planted += 1
        return reads.reshape(-1), np.arange(n + 1, dtype=np.uint64) * 150, planted
    lens = rng.integers(500, 5001, n)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(off[-1]))].copy()
    for r in np.nonzero(rng.random(n) < 0.3)[0]:
        for _ in range(int(rng.integers(1, 3))):
            h = hmms[int(rng.integers(n_hmm))]
            x = int(rng.integers(1, max(2, h.M // 2)))
            y = int(rng.integers(min(h.M, x + 30), h.M + 1))
            frag = np.frombuffer(back_translate(sample_protein(h, rng, x, y)).encode(), dtype=np.uint8)
            if len(frag) < lens[r]:
                p = int(off[r]) + int(rng.integers(0, lens[r] - len(frag)))
                nt[p:p + len(frag)] = frag
                planted += 1
    return nt, off, planted
