"""ctypes binding of libphyloaln_b200.so (C ABI: include/phyloaln_b200.h).

Thin by design: no torch, no numpy-side compute, no CPU fallback.  Importing this module without
the built CUDA library raises ImportError; creating a context without a B200 raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libphyloaln_b200.so")


class PaOpts(C.Structure):
    _fields_ = [("F1", C.c_double), ("F2", C.c_double), ("F3", C.c_double), ("do_bias", C.c_int),
                ("do_null2", C.c_int), ("do_max", C.c_int), ("reserved", C.c_int)]


class PaHit(C.Structure):
    _fields_ = [("hmm", C.c_int32), ("target", C.c_int32), ("ndom_in_seq", C.c_int32), ("dom_idx", C.c_int32),
                ("ienv", C.c_int32), ("jenv", C.c_int32), ("iali", C.c_int32), ("jali", C.c_int32),
                ("hmmfrom", C.c_int32), ("hmmto", C.c_int32), ("envsc", C.c_float), ("domcorrection", C.c_float),
                ("oasc", C.c_float), ("dombias", C.c_float), ("dom_bits", C.c_float), ("dom_lnP", C.c_double),
                ("seq_bits", C.c_float), ("pre_bits", C.c_float), ("seqbias", C.c_float), ("fwdsc", C.c_float),
                ("nullsc", C.c_float), ("seq_lnP", C.c_double), ("ali_off", C.c_int32), ("ali_len", C.c_int32),
                ("multidomain_region", C.c_int32)]


class PaStats(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("n_cells", C.c_uint64), ("n_past_ssv", C.c_uint64),
                ("n_past_msv", C.c_uint64), ("n_past_bias", C.c_uint64), ("n_past_vit", C.c_uint64),
                ("n_past_fwd", C.c_uint64), ("n_domains", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_translate", C.c_float), ("ms_msv", C.c_float), ("ms_bias", C.c_float), ("ms_vit", C.c_float),
                ("ms_fwd", C.c_float), ("ms_domain", C.c_float), ("ms_h2d", C.c_float), ("ms_total", C.c_float),
                ("ms_domain_kernel", C.c_float), ("vit_exact_pairs", C.c_float)]


class PaMapRec(C.Structure):
    _fields_ = [("keep", C.c_int32), ("qstart", C.c_int32), ("qend", C.c_int32), ("ali_from", C.c_int32),
                ("ali_to", C.c_int32), ("start_trim", C.c_int32), ("end_trim", C.c_int32), ("read_from", C.c_int32),
                ("read_to", C.c_int32), ("strand", C.c_int32), ("col_off", C.c_int32), ("ncols", C.c_int32),
                ("unmap_off", C.c_int32), ("n_unmap", C.c_int32), ("ext_start", C.c_int32), ("ext_end", C.c_int32)]


class PaMapOpts(C.Structure):
    _fields_ = [("trim_pos", C.c_int32), ("extend_pos", C.c_int32), ("gencode", C.c_int32), ("prefilter", C.c_int32),
                ("pos_freq", C.c_double), ("outgroup_weight", C.c_double)]


class PaAccum(C.Structure):
    _fields_ = [("bin", C.c_void_p), ("rep", C.c_void_p), ("nbins", C.c_int32), ("bin_row", C.c_void_p),
                ("hist", C.c_void_p), ("first", C.c_void_p), ("qrange", C.c_void_p)]


HIT_DTYPE = np.dtype(PaHit)
MAPREC_DTYPE = np.dtype(PaMapRec)

EXPORTS = ["pa_abi_version", "pa_create", "pa_destroy", "pa_last_error", "pa_set_opts", "pa_add_hmm", "pa_set_evparam",
           "pa_commit_hmms", "pa_commit_hmms_multi", "pa_num_hmms", "pa_load_reads", "pa_load_proteins", "pa_num_frames", "pa_get_targets",
           "pa_search", "pa_get_hits", "pa_get_hit_pp", "pa_hit_pool_bytes", "pa_get_stats", "pa_score_all", "pa_map_hits", "pa_map_hits_ex", "pa_merge_hists", "pa_microbench_dpx",
           "pa_profiler_range", "pa_msv_item_decode", "pa_hmm_file_parse",
           "pa_ingest_create", "pa_ingest_destroy", "pa_ingest_last_error", "pa_ingest_add_file", "pa_ingest_run", "pa_ingest_start",
           "pa_ingest_wait", "pa_ingest_done", "pa_ingest_emitted", "pa_ingest_num_records", "pa_ingest_views", "pa_ingest_write",
           "pa_ingest_pack", "pa_ingest_pack_emitted", "pa_ingest_emitted_bytes", "pa_ingest_emitted_ids", "pa_ingest_emitted_seqs", "pa_ingest_repeated_ids",
           "pa_ingest_final_order", "pa_fasta_open", "pa_fasta_num_records", "pa_fasta_views", "pa_fasta_close"]

_lib = None


def load_library():
    """dlopen the CUDA library; the product has no other execution path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with phyloaln_b200/csrc/build.sh "
                          "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64
    L.pa_abi_version.restype = i32
    L.pa_create.restype = vp
    L.pa_create.argtypes = [i32, C.c_char_p, i32]
    L.pa_destroy.argtypes = [vp]
    L.pa_last_error.restype = C.c_char_p
    L.pa_last_error.argtypes = [vp]
    L.pa_set_opts.argtypes = [vp, C.POINTER(PaOpts)]
    L.pa_add_hmm.argtypes = [vp, C.c_char_p, i32, i32, vp, vp, vp, vp, C.c_char_p, vp]
    L.pa_set_evparam.argtypes = [vp, i32, vp]
    L.pa_commit_hmms.argtypes = [vp]
    L.pa_commit_hmms_multi.argtypes = [vp, i32]
    L.pa_num_hmms.argtypes = [vp]
    L.pa_load_reads.restype = i64
    L.pa_load_reads.argtypes = [vp, vp, vp, u32, i32, i32, i32, i32]
    L.pa_load_proteins.restype = i64
    L.pa_load_proteins.argtypes = [vp, vp, vp, u32, i32]
    L.pa_num_frames.argtypes = [vp]
    L.pa_get_targets.restype = i64
    L.pa_get_targets.argtypes = [vp, vp, u64, vp]
    L.pa_search.restype = i64
    L.pa_search.argtypes = [vp]
    L.pa_get_hits.restype = i64
    L.pa_get_hits.argtypes = [vp, vp, i64, vp, vp, i64]
    L.pa_get_hit_pp.restype = i64
    L.pa_get_hit_pp.argtypes = [vp, vp, i64]
    L.pa_hit_pool_bytes.restype = i64
    L.pa_hit_pool_bytes.argtypes = [vp]
    L.pa_get_stats.argtypes = [vp, C.POINTER(PaStats)]
    L.pa_score_all.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.pa_map_hits.argtypes = [vp, C.c_int32, C.c_char_p, C.c_char_p, vp, C.c_char_p] + [vp] * 10 + [i32, C.c_double, vp, vp, vp, vp, i64, i64]
    L.pa_map_hits_ex.argtypes = ([vp, C.c_int32, C.c_char_p, C.c_char_p, vp, C.c_char_p] + [vp] * 13 +
                                 [C.POINTER(PaMapOpts), C.POINTER(PaAccum), vp, vp, vp, vp, i64, i64])
    L.pa_merge_hists.argtypes = [vp, C.c_int32] + [vp] * 11
    L.pa_microbench_dpx.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.pa_profiler_range.argtypes = [i32]
    _lib = L
    return L


def profiler_range(on: bool) -> None:
    """cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
    load_library().pa_profiler_range(1 if on else 0)


class PaError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data


def commit_multi(ctxs) -> None:
    """Commit the profiles added to ctxs[0] on every context of the list (configured and laid out once, uploaded to each GPU)."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = ctxs[0].L.pa_commit_hmms_multi(arr, len(ctxs))
    ctxs[0]._check(rc, "pa_commit_hmms_multi")
    for c in ctxs[1:]:
        c.names = list(ctxs[0].names)


def prepare_hmm(hmm, require_stats: bool = True):
    """The arguments of pa_add_hmm for one hmmio.HMM (contiguous float64 arrays, encoded strings), made once per profile."""
    mat = np.ascontiguousarray(hmm.mat, dtype=np.float64)
    ins = np.ascontiguousarray(hmm.ins, dtype=np.float64)
    t = np.ascontiguousarray(hmm.t, dtype=np.float64)
    compo = None if hmm.compo is None else np.ascontiguousarray(hmm.compo, dtype=np.float64)
    has = all(k in hmm.stats for k in ("MSV", "VITERBI", "FORWARD"))
    if require_stats and not has:
        hmm.evparams()  # raises the descriptive error
    ev = np.ascontiguousarray(hmm.evparams(), dtype=np.float32) if has else np.array([-10, 0.7, -11, 0.7, -4, 0.7], dtype=np.float32)
    return (hmm.name.encode(), 0 if hmm.abc == "amino" else 1, hmm.M, mat, ins, t, compo, hmm.consensus().encode(), ev)


class Context:
    """One search context bound to one GPU."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        err = C.create_string_buffer(512)
        self.h = self.L.pa_create(device, err, 512)
        if not self.h:
            raise PaError(f"pa_create failed: {err.value.decode()}")
        self.device = device
        self.names: list[str] = []

    def close(self):
        if getattr(self, "h", None):
            self.L.pa_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise PaError(f"{what} failed ({rc}): {self.L.pa_last_error(self.h).decode()}")
        return rc

    def set_opts(self, F1=0.02, F2=1e-3, F3=1e-5, do_bias=True, do_null2=True, do_max=False):
        o = PaOpts(F1, F2, F3, int(do_bias), int(do_null2), int(do_max), 0)
        self._check(self.L.pa_set_opts(self.h, C.byref(o)), "pa_set_opts")

    def add_hmm(self, hmm, require_stats: bool = True) -> int:
        """hmm: phyloaln_b200.hmmio.HMM, or the tuple prepare_hmm() made of it (several contexts share one conversion)"""
        name, abc, M, mat, ins, t, compo, cons, ev = hmm if isinstance(hmm, tuple) else prepare_hmm(hmm, require_stats)
        idx = self._check(self.L.pa_add_hmm(self.h, name, abc, M, _ptr(mat), _ptr(ins), _ptr(t), _ptr(compo), cons, _ptr(ev)),
                          "pa_add_hmm")
        self.names.append(name.decode())
        return idx

    def set_evparam(self, idx: int, ev):
        ev = np.ascontiguousarray(ev, dtype=np.float32)
        self._check(self.L.pa_set_evparam(self.h, idx, _ptr(ev)), "pa_set_evparam")

    def commit(self):
        self._check(self.L.pa_commit_hmms(self.h), "pa_commit_hmms")

    def load_reads(self, nt: np.ndarray, off: np.ndarray, moltype: int = 0, gencode: int = 1,
                   no_reverse: bool = False, codon_only: bool = False) -> int:
        """nt: uint8 array of concatenated ASCII nucleotides; off: uint64[nreads+1]."""
        nt = np.ascontiguousarray(nt, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        return self._check(self.L.pa_load_reads(self.h, _ptr(nt), _ptr(off), len(off) - 1, moltype, gencode,
                                                int(no_reverse), int(codon_only)), "pa_load_reads")

    def load_read_list(self, reads: list[str], **kw) -> int:
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r) for r in reads])
        nt = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
        if len(nt) == 0:
            nt = np.zeros(1, dtype=np.uint8)
        return self.load_reads(nt, off, **kw)

    def load_proteins(self, seqs: list[str], abc: int = 0) -> int:
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        aa = np.frombuffer("".join(seqs).encode(), dtype=np.uint8)
        if len(aa) == 0:
            aa = np.zeros(1, dtype=np.uint8)
        aa = np.ascontiguousarray(aa)
        return self._check(self.L.pa_load_proteins(self.h, _ptr(aa), _ptr(off), len(seqs), abc), "pa_load_proteins")

    def num_frames(self) -> int:
        return self.L.pa_num_frames(self.h)

    def get_targets(self, ntargets: int, cap: int) -> list[str]:
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        off = np.zeros(ntargets + 1, dtype=np.uint64)
        self._check(self.L.pa_get_targets(self.h, _ptr(out), cap, _ptr(off)), "pa_get_targets")
        b = out.tobytes()
        return [b[int(off[i]):int(off[i + 1])].decode() for i in range(ntargets)]

    def search(self):
        """Run the whole pipeline; returns (hits structured ndarray, model strings, aligned strings)."""
        n = self._check(self.L.pa_search(self.h), "pa_search")
        hits = np.zeros(n, dtype=HIT_DTYPE)
        pool = max(1, self.L.pa_hit_pool_bytes(self.h))
        mp, tp = C.create_string_buffer(pool), C.create_string_buffer(pool)
        self._check(self.L.pa_get_hits(self.h, _ptr(hits), n, mp, tp, pool), "pa_get_hits")
        mraw, traw = mp.raw, tp.raw          # one copy of each pool, then slices
        model = [mraw[o:o + l].decode() for o, l in zip(hits["ali_off"], hits["ali_len"])]
        aligned = [traw[o:o + l].decode() for o, l in zip(hits["ali_off"], hits["ali_len"])]
        return hits, model, aligned

    def hit_pp(self, hits):
        """PP annotation line of every alignment of the last search() (same order as its hits)."""
        pool = max(1, self.L.pa_hit_pool_bytes(self.h))
        pp = C.create_string_buffer(pool)
        self._check(self.L.pa_get_hit_pp(self.h, pp, pool), "pa_get_hit_pp")
        raw = pp.raw
        return [raw[o:o + l].decode() for o, l in zip(hits["ali_off"], hits["ali_len"])]

    def stats(self) -> dict:
        st = PaStats()
        self._check(self.L.pa_get_stats(self.h, C.byref(st)), "pa_get_stats")
        return {f: getattr(st, f) for f, _ in PaStats._fields_}

    def microbench_dpx(self, mode: int) -> float:
        v = C.c_double()
        self._check(self.L.pa_microbench_dpx(self.h, mode, C.byref(v)), "pa_microbench_dpx")
        return v.value

    def score_all(self, hmm_idx: int, ntargets: int, want=("msv", "vit", "fwd", "bias")):
        out = {}
        msv = np.zeros(ntargets, np.float32) if "msv" in want else None
        msv_raw = np.zeros(ntargets, np.int32) if "msv" in want else None
        vit = np.zeros(ntargets, np.float32) if "vit" in want else None
        vit_raw = np.zeros(ntargets, np.int32) if "vit" in want else None
        fwd = np.zeros(ntargets, np.float32) if "fwd" in want else None
        bias = np.zeros(ntargets, np.float32) if "bias" in want else None
        self._check(self.L.pa_score_all(self.h, hmm_idx, _ptr(msv), _ptr(msv_raw), _ptr(vit), _ptr(vit_raw), _ptr(fwd),
                                        _ptr(bias)), "pa_score_all")
        out.update(msv=msv, msv_raw=msv_raw, vit=vit, vit_raw=vit_raw, fwd=fwd, bias=bias)
        return out
