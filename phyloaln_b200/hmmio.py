"""HMMER3/f ASCII profile files: reader and writer (host side).

The reference builds one profile per reference alignment with ``hmmbuild`` and hands the
``.hmm`` file to ``hmmsearch`` (/root/reference/lib/aln.py:56-91, 664-667).  This module is the
host-side loader of those files for the B200 search path (format: SURVEY.md Appendix A.2).
Numbers are kept exactly as written in the file (negative natural logs, ``inf`` for ``*``) so
the C-ABI library performs the same ``expf(-x)`` conversion HMMER does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

AMINO = "ACDEFGHIKLMNPQRSTVWY"
DNA = "ACGT"
AMINO_SYMS = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
DNA_SYMS = "ACGT-RYMKSWHBVDN*~"
TRANS = ("MM", "MI", "MD", "IM", "II", "DM", "DD")


class HMMFormatError(ValueError):
    pass


@dataclass
class HMM:
    """One profile HMM. Arrays hold -ln(p); node 0 rows are the begin-node lines."""

    name: str
    abc: str                      # "amino" | "dna"
    mat: np.ndarray               # (M+1, K) float64, row 0 unused (inf)
    ins: np.ndarray               # (M+1, K)
    t: np.ndarray                 # (M+1, 7) MM MI MD IM II DM DD
    compo: np.ndarray | None = None
    cons: str | None = None       # M letters
    stats: dict = field(default_factory=dict)   # {"MSV": (mu, lam), "VITERBI": (mu, lam), "FORWARD": (tau, lam)}
    nseq: int = 1
    effn: float = 1.0
    rf: str | None = None
    map: list | None = None
    cutoffs: dict = field(default_factory=dict)  # {"GA": (per-seq bits, per-dom bits), "TC": ..., "NC": ...} when the file has them

    @property
    def M(self) -> int:
        return self.mat.shape[0] - 1

    @property
    def K(self) -> int:
        return self.mat.shape[1]

    @property
    def alphabet(self) -> str:
        return AMINO if self.abc == "amino" else DNA

    def evparams(self) -> np.ndarray:
        if not all(k in self.stats for k in ("MSV", "VITERBI", "FORWARD")):
            raise HMMFormatError(f"profile {self.name}: missing STATS LOCAL lines (uncalibrated)")
        return np.array([*self.stats["MSV"], *self.stats["VITERBI"], *self.stats["FORWARD"]], dtype=np.float32)

    def consensus(self) -> str:
        """Consensus line as HMMER's p7_hmm_SetConsensus would produce it when CONS is absent."""
        if self.cons is not None:
            return self.cons
        thr = 0.5 if self.abc == "amino" else 0.9
        p = np.exp(-self.mat[1:])
        out = []
        for k in range(self.M):
            b = int(np.argmax(p[k]))
            c = self.alphabet[b]
            out.append(c if p[k, b] >= thr else c.lower())
        return "".join(out)


def _val(tok: str) -> float:
    return math.inf if tok == "*" else float(tok)


def read_hmm(path: str) -> HMM:
    """One HMMER3/[a-f] ASCII profile (first model of the file) through the native reader (pa_hmm_file_parse, ~0.3 ms per
    profile instead of ~4 ms); a file it rejects is re-read by the Python statement below for the descriptive error."""
    import ctypes as C
    from . import capi
    L = capi.load_library()
    if not getattr(L, "_hmm_ready", False):
        L.pa_hmm_file_parse.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_char_p, C.c_int] + [C.c_void_p] * 8 + [
            C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L._hmm_ready = True
    M, abc, nseq, effn = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    name = C.create_string_buffer(512)
    stats = np.zeros(6, np.float32)
    if L.pa_hmm_file_parse(path.encode(), C.byref(M), C.byref(abc), name, 512, None, None, None, None, None, None, None,
                           stats.ctypes.data, C.byref(nseq), C.byref(effn)) != 0:
        return read_hmm_py(path)
    m, K = M.value, 20 if abc.value == 0 else 4
    mat, ins, t = np.empty((m + 1, K)), np.empty((m + 1, K)), np.empty((m + 1, 7))
    compo, mp = np.empty(K), np.zeros(max(m, 1), np.int32)
    cons, rf = C.create_string_buffer(m + 2), C.create_string_buffer(m + 2)
    if L.pa_hmm_file_parse(path.encode(), C.byref(M), C.byref(abc), name, 512, mat.ctypes.data, ins.ctypes.data, t.ctypes.data,
                           compo.ctypes.data, cons, rf, mp.ctypes.data, stats.ctypes.data, C.byref(nseq), C.byref(effn)) != 0:
        return read_hmm_py(path)
    st = {}
    for k, tag in enumerate(("MSV", "VITERBI", "FORWARD")):
        if not math.isnan(float(stats[2 * k])):
            st[tag] = (float(stats[2 * k]), float(stats[2 * k + 1]))
    return HMM(name=name.value.decode(), abc="amino" if abc.value == 0 else "dna", mat=mat, ins=ins, t=t,
               compo=None if np.isnan(compo).any() else compo, cons=cons.value.decode() or None, stats=st, nseq=nseq.value,
               effn=effn.value, rf=rf.value.decode() or None, map=None if mp[0] < 0 else [int(x) for x in mp[:m]],
               cutoffs=read_cutoffs(path))


def _cutoff_line(tok):
    """'GA 25.00 20.00;' -> (25.0, 20.0): Pfam-style per-sequence and per-domain bit-score thresholds."""
    return float(tok[1].rstrip(";")), float(tok[2].rstrip(";"))


def read_cutoffs(path: str) -> dict:
    """GA / TC / NC header lines of the first model of an HMMER3 ASCII file (the native reader does not return them)."""
    out = {}
    with open(path, "rb") as fh:          # bytes: a DESC line in some other encoding must not stop the header scan
        for raw in fh:
            if raw.startswith(b"HMM "):
                break
            if raw[:2] not in (b"GA", b"TC", b"NC"):
                continue
            tok = raw.decode("latin-1").split()
            if len(tok) >= 3 and tok[0] in ("GA", "TC", "NC"):
                try:
                    out[tok[0]] = _cutoff_line(tok)
                except ValueError:
                    raise HMMFormatError(f"{path}: bad {tok[0]} line") from None
    return out


def read_hmm_py(path: str) -> HMM:
    """Parse one HMMER3/[a-f] ASCII profile (first model of the file): the plain-Python statement of the format."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    if not lines or not lines[0].startswith("HMMER3/"):
        raise HMMFormatError(f"{path}: not a HMMER3 ASCII file")
    hdr = {}
    stats = {}
    cutoffs = {}
    i = 1
    while i < len(lines) and not lines[i].startswith("HMM "):
        tok = lines[i].split()
        if len(tok) >= 2:
            if tok[0] == "STATS" and len(tok) >= 5:
                stats[tok[2]] = (float(tok[3]), float(tok[4]))
            elif tok[0] in ("GA", "TC", "NC") and len(tok) >= 3:
                try:
                    cutoffs[tok[0]] = _cutoff_line(tok)
                except ValueError:
                    raise HMMFormatError(f"{path}: bad {tok[0]} line") from None
            else:
                hdr[tok[0]] = tok[1]
        i += 1
    if i >= len(lines):
        raise HMMFormatError(f"{path}: no HMM section")
    try:
        M = int(hdr["LENG"])
        alph = hdr["ALPH"].lower()
    except KeyError as e:
        raise HMMFormatError(f"{path}: missing header tag {e}") from None
    if alph == "amino":
        abc = "amino"
    elif alph in ("dna", "rna"):
        abc = "dna"
    else:
        raise HMMFormatError(f"{path}: unsupported alphabet {alph}")
    K = 20 if abc == "amino" else 4
    i += 2  # HMM line + transition-label line
    compo = None
    tok = lines[i].split()
    if tok and tok[0] == "COMPO":
        compo = np.array([_val(x) for x in tok[1:K + 1]])
        i += 1
    mat = np.full((M + 1, K), math.inf)
    ins = np.full((M + 1, K), math.inf)
    t = np.full((M + 1, 7), math.inf)
    ins[0] = [_val(x) for x in lines[i].split()[:K]]
    t[0] = [_val(x) for x in lines[i + 1].split()[:7]]
    i += 2
    has_cons = hdr.get("CONS", "no").lower() == "yes"
    has_rf = hdr.get("RF", "no").lower() == "yes"
    has_map = hdr.get("MAP", "no").lower() == "yes"
    cons, rf, mp = [], [], []
    for k in range(1, M + 1):
        tok = lines[i].split()
        if int(tok[0]) != k or len(tok) < K + 1:
            raise HMMFormatError(f"{path}: bad match line for node {k}")
        mat[k] = [_val(x) for x in tok[1:K + 1]]
        ann = tok[K + 1:]
        if has_map and len(ann) > 0:
            mp.append(int(ann[0]) if ann[0] != "-" else 0)
        if has_cons and len(ann) > 1:
            cons.append(ann[1])
        if has_rf and len(ann) > 2:
            rf.append(ann[2])
        ins[k] = [_val(x) for x in lines[i + 1].split()[:K]]
        t[k] = [_val(x) for x in lines[i + 2].split()[:7]]
        i += 3
    return HMM(name=hdr.get("NAME", "unnamed"), abc=abc, mat=mat, ins=ins, t=t, compo=compo,
               cons="".join(cons) if len(cons) == M else None, stats=stats,
               nseq=int(hdr.get("NSEQ", 1)), effn=float(hdr.get("EFFN", 1.0)),
               rf="".join(rf) if len(rf) == M else None, map=mp if len(mp) == M else None, cutoffs=cutoffs)


def _fmt(v: float) -> str:
    return "      *" if math.isinf(v) else f"{abs(v) if abs(v) < 5e-6 else v:7.5f}"


def write_hmm(hmm: HMM, path: str) -> None:
    """Write a HMMER3/f file that hmmsearch (and read_hmm) can load."""
    M, K = hmm.M, hmm.K
    cons = hmm.consensus()
    out = ["HMMER3/f [phyloaln_b200 hmmbuild-lite]", f"NAME  {hmm.name}", f"LENG  {M}",
           f"ALPH  {'amino' if hmm.abc == 'amino' else 'DNA'}", f"RF    {'yes' if hmm.rf else 'no'}", "MM    no",
           "CONS  yes", "CS    no", f"MAP   {'yes' if hmm.map else 'no'}", f"NSEQ  {hmm.nseq}",
           f"EFFN  {hmm.effn:f}"]
    for key in ("GA", "TC", "NC"):
        if key in hmm.cutoffs:
            a, b = hmm.cutoffs[key]
            out.append(f"{key}    {a:.2f} {b:.2f};")
    for key in ("MSV", "VITERBI", "FORWARD"):
        if key in hmm.stats:
            a, b = hmm.stats[key]
            out.append(f"STATS LOCAL {key:<7s} {a:9.4f} {b:8.5f}")
    out.append("HMM     " + "".join(f"     {c}   " for c in hmm.alphabet))
    out.append("            m->m     m->i     m->d     i->m     i->i     d->m     d->d")
    if hmm.compo is not None:
        out.append("  COMPO   " + "  ".join(_fmt(v) for v in hmm.compo))
    out.append("          " + "  ".join(_fmt(v) for v in hmm.ins[0]))
    out.append("          " + "  ".join(_fmt(v) for v in hmm.t[0]))
    for k in range(1, M + 1):
        mapv = str(hmm.map[k - 1]) if hmm.map else "-"
        rfv = hmm.rf[k - 1] if hmm.rf else "-"
        out.append(f"{k:7d}   " + "  ".join(_fmt(v) for v in hmm.mat[k]) + f" {mapv:>6s} {cons[k - 1]} {rfv} - -")
        out.append("          " + "  ".join(_fmt(v) for v in hmm.ins[k]))
        out.append("          " + "  ".join(_fmt(v) for v in hmm.t[k]))
    out.append("//")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
