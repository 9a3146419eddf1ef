"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never imports it.
Parity status: UNPINNED at the HMMER boundary (see oracle/p7oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcOpts(C.Structure):
    _fields_ = [("F1", C.c_double), ("F2", C.c_double), ("F3", C.c_double), ("E", C.c_double),
                ("domE", C.c_double), ("do_bias", C.c_int), ("do_null2", C.c_int), ("do_max", C.c_int),
                ("Z", C.c_double), ("domZ", C.c_double), ("simd_msv", C.c_int), ("simd_vit", C.c_int)]


class OrcHit(C.Structure):
    _fields_ = [("target", C.c_int), ("ndom_in_seq", C.c_int), ("dom_idx", C.c_int), ("ienv", C.c_int),
                ("jenv", C.c_int), ("iali", C.c_int), ("jali", C.c_int), ("hmmfrom", C.c_int), ("hmmto", C.c_int),
                ("envsc", C.c_float), ("domcorrection", C.c_float), ("oasc", C.c_float), ("dombias", C.c_float),
                ("dom_bits", C.c_float), ("dom_lnP", C.c_double), ("seq_bits", C.c_float), ("pre_bits", C.c_float),
                ("seqbias", C.c_float), ("fwdsc", C.c_float), ("nullsc", C.c_float), ("seq_lnP", C.c_double),
                ("ali_off", C.c_int), ("ali_len", C.c_int), ("multidomain_region", C.c_int)]


class OrcTrace(C.Structure):
    _fields_ = [("nullsc", C.c_float), ("usc", C.c_float), ("filtersc", C.c_float), ("vfsc", C.c_float),
                ("fwdsc", C.c_float), ("msv_xJ", C.c_int), ("vit_xC", C.c_int), ("fwd_exp", C.c_int),
                ("fwd_mant", C.c_float), ("stage", C.c_int)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_hmm_read.restype = C.c_void_p
        L.orc_hmm_read.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.orc_hmm_from_arrays.restype = C.c_void_p
        L.orc_hmm_from_arrays.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_char_p, C.c_void_p, C.c_char_p]
        L.orc_hmm_free.argtypes = [C.c_void_p]
        L.orc_hmm_M.argtypes = [C.c_void_p]
        L.orc_hmm_abc.argtypes = [C.c_void_p]
        L.orc_hmm_set_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_hmm_name.restype = C.c_char_p
        L.orc_hmm_name.argtypes = [C.c_void_p]
        L.orc_profile_create.restype = C.c_void_p
        L.orc_profile_create.argtypes = [C.c_void_p]
        L.orc_profile_free.argtypes = [C.c_void_p]
        L.orc_digitize.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_void_p]
        L.orc_null1.restype = C.c_float
        L.orc_null1.argtypes = [C.c_int]
        for f in (L.orc_msv, L.orc_vit, L.orc_msv_simd, L.orc_vit_simd):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.orc_search_many.restype = C.c_longlong
        L.orc_search_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcOpts), C.c_int, C.c_void_p]
        L.orc_translate_reads.restype = C.c_longlong
        L.orc_translate_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_bias_filtersc.restype = C.c_float
        L.orc_bias_filtersc.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_fwd_parser.restype = C.c_float
        L.orc_fwd_parser.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_float)]
        L.orc_bwd_parser_total.restype = C.c_float
        L.orc_bwd_parser_total.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        for f in (L.orc_generic_forward, L.orc_generic_viterbi):
            f.restype = C.c_double
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcOpts), C.c_void_p,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_domain_decoding.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_translate.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.orc_revcomp.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        L.orc_codon_table.argtypes = [C.c_int, C.c_char_p]
        for f in (L.orc_gumbel_surv, L.orc_exp_surv, L.orc_exp_logsurv):
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_profile_tsc.restype = C.c_float
        L.orc_profile_tsc.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_profile_msc.restype = C.c_float
        L.orc_profile_msc.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in (L.orc_profile_rbv, L.orc_profile_rwv):
            f.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_profile_bytes.argtypes = [C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


AMINO_SYMS = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
DNA_SYMS = "ACGT-RYMKSWHBVDN*~"


def default_opts(**kw) -> OrcOpts:
    o = OrcOpts(0.02, 1e-3, 1e-5, 10.0, 10.0, 1, 1, 0, 0.0, 0.0, 0, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Profile:
    """One oracle profile (HMM + configured score tables)."""

    def __init__(self, hmm_handle):
        self.L = lib()
        self.h = hmm_handle
        self.p = self.L.orc_profile_create(self.h)
        self.M = self.L.orc_hmm_M(self.h)
        self.abc = self.L.orc_hmm_abc(self.h)
        self.name = self.L.orc_hmm_name(self.h).decode()

    @classmethod
    def from_file(cls, path: str) -> "Profile":
        err = C.create_string_buffer(512)
        h = lib().orc_hmm_read(path.encode(), err, 512)
        if not h:
            raise ValueError(err.value.decode())
        return cls(h)

    @classmethod
    def from_hmm(cls, hmm) -> "Profile":
        """hmm: phyloaln_b200.hmmio.HMM (arrays of -ln p)."""
        mat = np.ascontiguousarray(hmm.mat, dtype=np.float64)
        ins = np.ascontiguousarray(hmm.ins, dtype=np.float64)
        t = np.ascontiguousarray(hmm.t, dtype=np.float64)
        compo = None if hmm.compo is None else np.ascontiguousarray(hmm.compo, dtype=np.float64)
        ev = None
        if all(k in hmm.stats for k in ("MSV", "VITERBI", "FORWARD")):
            ev = np.ascontiguousarray(hmm.evparams(), dtype=np.float32)
        cons = hmm.consensus().encode()
        h = lib().orc_hmm_from_arrays(hmm.M, 0 if hmm.abc == "amino" else 1, mat.ctypes.data, ins.ctypes.data,
                                      t.ctypes.data, None if compo is None else compo.ctypes.data, cons,
                                      None if ev is None else ev.ctypes.data, hmm.name.encode())
        return cls(h)

    def set_stats(self, ev):
        ev = np.ascontiguousarray(ev, dtype=np.float32)
        self.L.orc_hmm_set_stats(self.h, ev.ctypes.data)

    def digitize(self, seq: str) -> np.ndarray:
        """1-based digital sequence with sentinels (dsq[0], dsq[L+1] = 255)."""
        b = seq.encode()
        out = np.full(len(b) + 2, 255, dtype=np.uint8)
        self.L.orc_digitize(self.abc, b, len(b), out[1:].ctypes.data)
        return out

    def msv(self, seq: str):
        d = self.digitize(seq)
        sc, raw = C.c_float(), C.c_int()
        self.L.orc_msv(self.p, d.ctypes.data, len(seq), C.byref(sc), C.byref(raw))
        return sc.value, raw.value

    def msv_simd(self, seq: str):
        """(score, raw xJ) from the striped AVX2 filter, or None without AVX2."""
        d = self.digitize(seq)
        sc, raw = C.c_float(), C.c_int()
        rc = self.L.orc_msv_simd(self.p, d.ctypes.data, len(seq), C.byref(sc), C.byref(raw))
        return None if rc < 0 else (sc.value, raw.value)

    def vit_simd(self, seq: str):
        """(score, raw xC) from the striped AVX2 Viterbi filter, or None without AVX2."""
        d = self.digitize(seq)
        sc, raw = C.c_float(), C.c_int()
        rc = self.L.orc_vit_simd(self.p, d.ctypes.data, len(seq), C.byref(sc), C.byref(raw))
        return None if rc < 0 else (sc.value, raw.value)

    def vit(self, seq: str):
        d = self.digitize(seq)
        sc, raw = C.c_float(), C.c_int()
        self.L.orc_vit(self.p, d.ctypes.data, len(seq), C.byref(sc), C.byref(raw))
        return sc.value, raw.value

    def bias(self, seq: str) -> float:
        d = self.digitize(seq)
        return self.L.orc_bias_filtersc(self.p, d.ctypes.data, len(seq))

    def fwd(self, seq: str, multihit: int = 1, Lcfg: int | None = None):
        d = self.digitize(seq)
        e, m = C.c_int(), C.c_float()
        sc = self.L.orc_fwd_parser(self.p, d.ctypes.data, len(seq), Lcfg or len(seq), multihit, C.byref(e), C.byref(m))
        return sc, e.value, m.value

    def bwd(self, seq: str, multihit: int = 1) -> float:
        d = self.digitize(seq)
        return self.L.orc_bwd_parser_total(self.p, d.ctypes.data, len(seq), len(seq), multihit)

    def generic_forward(self, seq: str, multihit: int = 1) -> float:
        d = self.digitize(seq)
        return self.L.orc_generic_forward(self.p, d.ctypes.data, len(seq), len(seq), multihit)

    def generic_viterbi(self, seq: str, multihit: int = 1) -> float:
        d = self.digitize(seq)
        return self.L.orc_generic_viterbi(self.p, d.ctypes.data, len(seq), len(seq), multihit)

    def domain_decoding(self, seq: str):
        d = self.digitize(seq)
        n = len(seq) + 1
        b, e, m = (np.zeros(n, dtype=np.float32) for _ in range(3))
        self.L.orc_domain_decoding(self.p, d.ctypes.data, len(seq), b.ctypes.data, e.ctypes.data, m.ctypes.data)
        return b, e, m

    def search(self, seqs: list[str], opts: OrcOpts | None = None, nthreads: int = 1, want_trace: bool = True):
        """Run the pipeline over ASCII targets. Returns (hits[list of dict], trace ndarray|None)."""
        opts = opts or default_opts()
        n = len(seqs)
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        cat = "".join(seqs).encode()
        dsq = np.zeros(max(1, len(cat)), dtype=np.uint8)
        self.L.orc_digitize(self.abc, cat, len(cat), dsq.ctypes.data)
        maxhits = max(1024, 4 * n)
        poolcap = max(1 << 20, 16 * len(cat) + 8 * self.M * 64)
        hits = (OrcHit * maxhits)()
        mp, tp = C.create_string_buffer(poolcap), C.create_string_buffer(poolcap)
        tr = (OrcTrace * max(1, n))() if want_trace else None
        nh = self.L.orc_search(self.p, dsq.ctypes.data, off.ctypes.data, n, C.byref(opts), hits, maxhits, mp, tp,
                               poolcap, tr, nthreads)
        if nh < 0:
            raise MemoryError("oracle hit buffers too small")
        out = []
        for i in range(nh):
            h = hits[i]
            d = {f: getattr(h, f) for f, _ in OrcHit._fields_}
            # string_at copies just the slice; `.raw` would copy the whole pool for every hit
            d["model"] = C.string_at(C.addressof(mp) + h.ali_off, h.ali_len).decode()
            d["aligned"] = C.string_at(C.addressof(tp) + h.ali_off, h.ali_len).decode()
            d["pp"] = C.string_at(C.addressof(tp) + h.ali_off + h.ali_len + 1, h.ali_len).decode()
            out.append(d)
        trace = None
        if want_trace:
            trace = np.ctypeslib.as_array(tr).copy() if n else None
        return out, trace

    def score_all(self, seqs: list[str], simd: bool = False):
        """Raw (msv, vit, fwd) nats for every sequence (no filtering) -- used for calibration.  simd: the striped AVX2
        filters (same numbers, tests/test_oracle.py) where the CPU has them."""
        fm = (lambda s: self.msv_simd(s) or self.msv(s)) if simd else self.msv
        fv = (lambda s: self.vit_simd(s) or self.vit(s)) if simd else self.vit
        msv = np.array([fm(s)[0] for s in seqs], dtype=np.float64)
        vit = np.array([fv(s)[0] for s in seqs], dtype=np.float64)
        fwd = np.array([self.fwd(s)[0] for s in seqs], dtype=np.float64)
        return msv, vit, fwd


def translate(nt: str, table: int = 1) -> str:
    b = nt.encode()
    out = C.create_string_buffer(len(b) // 3 + 1)
    n = lib().orc_translate(b, len(b), table, out)
    if n == -1:
        raise ValueError("invalid nucleotide character (Biopython would raise TranslationError)")
    if n == -2:
        raise ValueError(f"unknown genetic code table {table}")
    return out.raw[:n].decode()


def revcomp(nt: str) -> str:
    b = nt.encode()
    out = C.create_string_buffer(len(b) + 1)
    lib().orc_revcomp(b, len(b), out)
    return out.raw[:len(b)].decode()


def six_frames(nt: str, table: int = 1, no_reverse: bool = False, codon: bool = False):
    """(suffix, protein) records in the reference's order (lib/aln.py:249-258)."""
    out = []
    for j in (1, 2, 3):
        out.append((f"_pos{j}", translate(nt[j - 1:], table)))
        if codon:
            break
    if not no_reverse:
        rc = revcomp(nt)
        for j in (1, 2, 3):
            out.append((f"_rev_pos{j}", translate(rc[j - 1:], table)))
    return out


def translate_reads(nt: np.ndarray, off: np.ndarray, table: int = 1, nthreads: int = 1):
    """Six-frame translation + digitisation of concatenated reads in C: (dsq uint8, off int64[6n+1])."""
    nt = np.ascontiguousarray(nt, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = len(off) - 1
    dsq = np.zeros(2 * int(off[-1]) + 8, dtype=np.uint8)
    toff = np.zeros(6 * n + 1, dtype=np.int64)
    rc = lib().orc_translate_reads(nt.ctypes.data, off.ctypes.data, n, table, dsq.ctypes.data, toff.ctypes.data, nthreads)
    if rc < 0:
        raise ValueError("invalid nucleotide character (Biopython would raise TranslationError)")
    return dsq, toff


def search_many(profiles, dsq: np.ndarray, off: np.ndarray, opts: OrcOpts | None = None, nthreads: int = 1):
    """Every profile against every digital target in one OpenMP region (the CPU baseline of bench.py).
    Returns (number of domains, stage counts[6])."""
    opts = opts or default_opts()
    arr = (C.c_void_p * len(profiles))(*[p.p for p in profiles])
    st = np.zeros(6, dtype=np.int64)
    nd = lib().orc_search_many(arr, len(profiles), dsq.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(opts), nthreads,
                               st.ctypes.data)
    if nd < 0:
        raise MemoryError("oracle hit buffers too small")
    return int(nd), st
