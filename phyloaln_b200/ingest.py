"""Native target preparation -- the host mirror of the reference's ``prepare_target``
(/root/reference/lib/aln.py:274-433): FASTQ/FASTA(.gz) -> tagged, de-duplicated ``all.raw.fasta`` +
``total_counts.tsv`` + ``merged_redundance.txt``, byte-identical to the reference's files.

The work is done by ``pa_ingest_*`` in libphyloaln_b200.so (csrc/ingest.cpp: one inflate/parse thread per
input file, then the reference's order-dependent single pass).  ``Ingest.pack`` hands contiguous read
chunks straight to ``Context.load_reads`` so the GPU search never re-parses the FASTA text.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import capi


def _lib():
    L = capi.load_library()
    if not getattr(L, "_ingest_ready", False):
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        L.pa_ingest_create.restype = vp
        L.pa_ingest_create.argtypes = [i32, i32, i32, i32]
        L.pa_ingest_destroy.argtypes = [vp]
        L.pa_ingest_last_error.restype = C.c_char_p
        L.pa_ingest_last_error.argtypes = [vp]
        L.pa_ingest_add_file.argtypes = [vp, C.c_char_p, C.c_char_p, i32, C.c_char_p]
        L.pa_ingest_run.argtypes = [vp]
        L.pa_ingest_num_records.restype = i64
        L.pa_ingest_num_records.argtypes = [vp]
        L.pa_ingest_views.argtypes = [vp] + [vp] * 10
        L.pa_ingest_write.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p]
        L.pa_ingest_pack.restype = i64
        L.pa_ingest_pack.argtypes = [vp, u64, u64, vp, u64, vp]
        for f in ("pa_ingest_start", "pa_ingest_wait", "pa_ingest_done"):
            getattr(L, f).argtypes = [vp]
        L.pa_ingest_emitted.restype = i64
        L.pa_ingest_emitted.argtypes = [vp]
        L.pa_ingest_pack_emitted.restype = i64
        L.pa_ingest_pack_emitted.argtypes = [vp, u64, u64, vp, u64, vp]
        L.pa_ingest_emitted_bytes.restype = i64
        L.pa_ingest_emitted_bytes.argtypes = [vp, u64, u64]
        L.pa_ingest_emitted_ids.restype = i64
        L.pa_ingest_emitted_ids.argtypes = [vp, vp, u64, vp, u64, vp]
        L.pa_ingest_emitted_seqs.restype = i64
        L.pa_ingest_emitted_seqs.argtypes = [vp, vp, u64, vp, u64, vp]
        L.pa_ingest_repeated_ids.restype = i64
        L.pa_ingest_repeated_ids.argtypes = [vp]
        L.pa_ingest_final_order.argtypes = [vp, vp]
        L.pa_fasta_open.restype = vp
        L.pa_fasta_open.argtypes = [C.c_char_p, C.c_char_p, i32]
        L.pa_fasta_num_records.restype = i64
        L.pa_fasta_num_records.argtypes = [vp]
        L.pa_fasta_views.argtypes = [vp] + [vp] * 4
        L.pa_fasta_close.argtypes = [vp]
        L._ingest_ready = True
    return L


class LazyIds:
    """Sequence of record ids decoded on demand from one byte blob (10 M ids are never all needed: only reads with hits)."""

    def __init__(self, blob, off: np.ndarray):     # blob: bytes or a uint8 array
        self._blob, self._off = blob, off

    def __len__(self):
        return len(self._off) - 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return bytes(self._blob[int(self._off[i]):int(self._off[i + 1])]).decode()

    def __iter__(self):
        return (self[k] for k in range(len(self)))


def load_raw_fasta(path: str):
    """Native all.raw.fasta loader (pa_fasta_*): (ids, nt uint8 array, offsets uint64[n+1]) with the dictionary semantics of
    the reference's read_fastx (lib/library.py:186-198); ids is a lazily decoding sequence."""
    L = _lib()
    err = C.create_string_buffer(512)
    h = L.pa_fasta_open(path.encode(), err, 512)
    if not h:
        raise OSError(err.value.decode() or f"cannot read {path}")
    holder = _FastaHandle(L, h)
    n = L.pa_fasta_num_records(h)
    p_ids, p_idoff, p_nt, p_off = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    L.pa_fasta_views(h, C.byref(p_ids), C.byref(p_idoff), C.byref(p_nt), C.byref(p_off))

    def view(ptr, ctype, count, dtype):      # zero-copy: the array keeps the native handle alive through its base
        if count == 0 or not ptr.value:
            return np.zeros(0, dtype=dtype)
        raw = (ctype * count).from_address(ptr.value)
        raw._holder = holder
        return np.frombuffer(raw, dtype=dtype)
    id_off = view(p_idoff, C.c_uint64, n + 1, np.uint64)
    off = view(p_off, C.c_uint64, n + 1, np.uint64)
    if n == 0:
        return LazyIds(b"", np.zeros(1, np.uint64)), np.zeros(0, np.uint8), np.zeros(1, np.uint64)
    blob = view(p_ids, C.c_uint8, int(id_off[-1]), np.uint8)
    nt = view(p_nt, C.c_uint8, int(off[-1]), np.uint8)
    return LazyIds(blob, id_off), nt, off


class _FastaHandle:
    """Owner of a pa_fasta: closed when the last array viewing its buffers is gone."""

    def __init__(self, L, h):
        self.L, self.h = L, h

    def __del__(self):
        if self.h:
            self.L.pa_fasta_close(self.h)
            self.h = None


class Ingest:
    """rawdata: {species: [files]} in the reference's iteration order (dict order, lib/aln.py:287-289)."""

    def __init__(self, rawdata: dict, merge_len: int = 250, split_len: int | None = None, split_slide: int | None = None,
                 file_format: str = "guess", threads: int | None = None, wait: bool = True):
        """wait=False returns while the files are still being decoded: records stream out in emission order
        (emitted / pack_emitted / ids_of) and the search can start on them; call wait() before anything else."""
        self.L = _lib()
        if split_len is not None and split_slide is None:
            split_slide = int(split_len / 2)          # lib/main.py:207-208
        self.h = self.L.pa_ingest_create(int(merge_len), int(split_len or 0), int(split_slide or 0),
                                         int(threads or os.cpu_count() or 1))
        for sp, files in rawdata.items():
            for j, f in enumerate(files):
                self._check(self.L.pa_ingest_add_file(self.h, os.fspath(f).encode(), sp.encode(), j + 1, file_format.encode()))
        self.n = -1
        self._wait_lock = threading.Lock()
        self._check(self.L.pa_ingest_start(self.h))
        if wait:
            self.wait()

    def wait(self):
        """Join the background pass; raises on its error. Idempotent."""
        with self._wait_lock:
            if self.n < 0:
                self._check(self.L.pa_ingest_wait(self.h))
                self.n = int(self.L.pa_ingest_num_records(self.h))
        return self

    def done(self) -> bool:
        return bool(self.L.pa_ingest_done(self.h))

    def emitted(self) -> int:
        return int(self.L.pa_ingest_emitted(self.h))

    def emitted_bytes(self, first: int, count: int) -> int:
        return int(self.L.pa_ingest_emitted_bytes(self.h, first, count))

    def pack_emitted(self, first: int, count: int):
        """(nt uint8[], off uint64[count+1]) of emitted records [first, first+count); valid while the ingest runs."""
        cap = self.emitted_bytes(first, count) + 1
        nt = np.zeros(cap, dtype=np.uint8)
        off = np.zeros(count + 1, dtype=np.uint64)
        got = self.L.pa_ingest_pack_emitted(self.h, first, count, nt.ctypes.data, cap, off.ctypes.data)
        self._check(got)
        return nt[:max(int(got), 1)], off

    def _fields_of(self, fn, idx) -> list[str]:
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        off = np.zeros(len(idx) + 1, dtype=np.uint64)
        need = fn(self.h, idx.ctypes.data, len(idx), None, 0, off.ctypes.data)
        self._check(need)
        buf = C.create_string_buffer(max(1, int(need)))
        self._check(fn(self.h, idx.ctypes.data, len(idx), buf, int(need), off.ctypes.data))
        raw = buf.raw
        return [raw[int(off[i]):int(off[i + 1])].decode() for i in range(len(idx))]

    def ids_of(self, idx) -> list[str]:
        """Tagged ids of the records with emission indices idx (only reads with hits ever need their id as a str)."""
        return self._fields_of(self.L.pa_ingest_emitted_ids, idx)

    def seqs_of(self, idx) -> list[str]:
        """Raw nucleotide sequences of the records with emission indices idx (targets.fas, lib/aln.py:559-561)."""
        return self._fields_of(self.L.pa_ingest_emitted_seqs, idx)

    def repeated_ids(self) -> int:
        self.wait()
        return int(self.L.pa_ingest_repeated_ids(self.h))

    def final_order(self):
        """None when all.raw.fasta order == emission order, else emission index of every final position."""
        self.wait()
        perm = np.zeros(max(1, self.n), dtype=np.uint64)
        rc = self.L.pa_ingest_final_order(self.h, perm.ctypes.data)
        self._check(rc)
        return perm[:self.n] if rc == 1 else None

    def _check(self, rc):
        if rc < 0:
            raise capi.PaError(f"ingest failed ({rc}): {self.L.pa_ingest_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.pa_ingest_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _views(self):
        self.wait()
        raw, cnt, mrg = C.c_char_p(), C.c_char_p(), C.c_char_p()
        rl, cl, ml = C.c_uint64(), C.c_uint64(), C.c_uint64()
        ptrs = [C.POINTER(C.c_uint64)() for _ in range(4)]
        raw_p, cnt_p, mrg_p = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.L.pa_ingest_views(self.h, C.byref(raw_p), C.byref(rl), C.byref(ptrs[0]), C.byref(ptrs[1]),
                                           C.byref(ptrs[2]), C.byref(ptrs[3]), C.byref(cnt_p), C.byref(cl), C.byref(mrg_p),
                                           C.byref(ml)))
        del raw, cnt, mrg
        return raw_p, rl.value, ptrs, cnt_p, cl.value, mrg_p, ml.value

    def raw_fasta(self) -> bytes:
        raw_p, rl, *_ = self._views()
        return C.string_at(raw_p, rl) if rl else b""

    def total_counts(self) -> str:
        _, _, _, cnt_p, cl, _, _ = self._views()
        return C.string_at(cnt_p, cl).decode() if cl else ""

    def merged_redundance(self) -> str:
        *_, mrg_p, ml = self._views()
        return C.string_at(mrg_p, ml).decode() if ml else ""

    def ids(self) -> list[str]:
        raw_p, rl, ptrs, *_ = self._views()
        raw = C.string_at(raw_p, rl) if rl else b""
        io = np.ctypeslib.as_array(ptrs[0], shape=(self.n,)) if self.n else np.zeros(0, np.uint64)
        il = np.ctypeslib.as_array(ptrs[1], shape=(self.n,)) if self.n else np.zeros(0, np.uint64)
        return [raw[int(o):int(o + l)].decode() for o, l in zip(io, il)]

    def seq_lengths(self) -> np.ndarray:
        _, _, ptrs, *_ = self._views()
        return np.ctypeslib.as_array(ptrs[3], shape=(self.n,)).copy() if self.n else np.zeros(0, np.uint64)

    def pack(self, first: int = 0, count: int | None = None):
        """(nt uint8[], off uint64[count+1]) of records [first, first+count) for ``Context.load_reads``."""
        self.wait()
        count = self.n - first if count is None else count
        cap = int(self.seq_lengths()[first:first + count].sum()) + 1
        nt = np.zeros(cap, dtype=np.uint8)
        off = np.zeros(count + 1, dtype=np.uint64)
        got = self.L.pa_ingest_pack(self.h, first, count, nt.ctypes.data, cap, off.ctypes.data)
        self._check(got)
        return nt[:max(int(got), 1)], off

    def write(self, outdir: str = "."):
        """all.raw.fasta, total_counts.tsv, merged_redundance.txt exactly where the reference puts them."""
        self.wait()
        self._check(self.L.pa_ingest_write(self.h, os.path.join(outdir, "all.raw.fasta").encode(),
                                           os.path.join(outdir, "total_counts.tsv").encode(),
                                           os.path.join(outdir, "merged_redundance.txt").encode()))


def prepare_target(rawdata: dict, args, outdir: str = ".", wait: bool = True) -> Ingest:
    """Drop-in for the ingest half of the reference's prepare_target(rawdata, args): same arguments
    (args.file_format, args.split_len, args.split_slide, args.merge_len, args.cpu), same three files.
    wait=False: return while the files are still being decoded (the caller writes the files after wait())."""
    ing = Ingest(rawdata, merge_len=args.merge_len, split_len=args.split_len, split_slide=args.split_slide,
                 file_format=args.file_format, threads=getattr(args, "cpu", None), wait=wait)
    if wait:
        ing.write(outdir)
    return ing
