"""Host side of the B200 search step: option parsing, sharding, report thresholds and the
PhyloAln-facing outputs.

Mirrors, for the default ``-j hmmer`` mode, what the reference does around the ``hmmsearch``
child process (/root/reference/lib/aln.py):

* ``assemble()`` :650-674 builds ``hmmsearch -o .. --tblout .. --cpu N <params> g.hmm all.target.fasta``;
  here :func:`parse_hmmsearch_parameters` understands the same pass-through parameter list and
  :class:`B200Searcher` runs all profiles against a resident read chunk on the GPU.
* HMMER applies the E-value thresholds with ``Z`` = number of target sequences and ``domZ`` =
  number of reported sequences after all targets are seen and sorts the report
  (A.10); :func:`report` does that on the gathered hit records (the only global step, host side,
  no collective on the data path).
* ``extract_reads()`` :482-570 turns the text report into ``map/{g}.hit_info.tsv`` rows and
  ``map/{g}.targets.fas``; :func:`hit_rows` / :func:`write_step_outputs` produce the same files
  and ``ok/`` markers directly from the hit records.
"""
from __future__ import annotations

import math
import os
import threading
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .hmmio import HMM, read_hmm


class SearchParameterError(ValueError):
    pass


@dataclass
class SearchOptions:
    """hmmsearch options that affect the hit set (reference: args.hmmsearch_parameters, lib/aln.py:878)."""
    E: float = 10.0
    domE: float = 10.0
    T: float | None = None
    domT: float | None = None
    incE: float = 0.01
    incdomE: float = 0.01
    Z: float | None = None
    domZ: float | None = None
    F1: float = 0.02
    F2: float = 1e-3
    F3: float = 1e-5
    nobias: bool = False
    nonull2: bool = False
    max: bool = False
    cut: str | None = None        # "GA" | "TC" | "NC": --cut_ga / --cut_tc / --cut_nc, the profile's own bit-score thresholds
    cutoffs: np.ndarray | None = None   # (n_hmm, 2) per-sequence / per-domain bits, filled by resolve_cutoffs()
    ignored: list = field(default_factory=list)


_FLOAT_OPTS = {"-E": "E", "--domE": "domE", "-T": "T", "--domT": "domT", "--incE": "incE", "--incdomE": "incdomE",
               "-Z": "Z", "--domZ": "domZ", "--F1": "F1", "--F2": "F2", "--F3": "F3"}
_FLAG_OPTS = {"--nobias": "nobias", "--nonull2": "nonull2", "--max": "max"}
_IGNORED_WITH_ARG = {"--cpu", "--seed", "-o", "-A", "--tblout", "--domtblout", "--pfamtblout", "--textw", "--tformat"}
_IGNORED_FLAGS = {"--acc", "--noali", "--notextw"}
_CUT_OPTS = {"--cut_ga": "GA", "--cut_tc": "TC", "--cut_nc": "NC"}
_CUT_EXCLUDES = ("-E", "-T", "--domE", "--domT", "--incE", "--incT", "--incdomE", "--incdomT")   # as HMMER's option table


def parse_hmmsearch_parameters(params) -> SearchOptions:
    """Parse PhyloAln's pass-through list (e.g. [' -E', '1e-5']; leading blanks as in lib/main.py:11-12)."""
    o = SearchOptions()
    toks = []
    for p in params or []:
        toks.extend(str(p).split())
    i = 0
    while i < len(toks):
        t = toks[i]
        if t in _FLOAT_OPTS:
            if i + 1 >= len(toks):
                raise SearchParameterError(f"hmmsearch option {t} needs a value")
            try:
                setattr(o, _FLOAT_OPTS[t], float(toks[i + 1]))
            except ValueError:
                raise SearchParameterError(f"bad value for {t}: {toks[i + 1]}") from None
            i += 2
        elif t in _FLAG_OPTS:
            setattr(o, _FLAG_OPTS[t], True)
            i += 1
        elif t in _IGNORED_WITH_ARG:
            o.ignored.append(t)
            i += 2
        elif t in _CUT_OPTS:
            if o.cut is not None and o.cut != _CUT_OPTS[t]:
                raise SearchParameterError(f"{t} is incompatible with --cut_{o.cut.lower()}")
            o.cut = _CUT_OPTS[t]
            i += 1
        elif t in _IGNORED_FLAGS:
            o.ignored.append(t)
            i += 1
        else:
            raise SearchParameterError(f"unsupported hmmsearch option for the B200 search path: {t}")
    if o.cut is not None:
        for t in toks:
            if t in _CUT_EXCLUDES:
                raise SearchParameterError(f"--cut_{o.cut.lower()} is incompatible with option {t}")
    return o


def resolve_cutoffs(opts: SearchOptions, hmms) -> None:
    """--cut_ga / --cut_tc / --cut_nc: take every profile's own thresholds (HMMER's p7_pli_NewModelThresholds: per-sequence
    score >= GA1, per-domain score >= GA2, E-values no longer decide).  A profile without them is an error, as in HMMER."""
    if opts.cut is None:
        opts.cutoffs = None
        return
    c = np.empty((len(hmms), 2), dtype=np.float64)
    for i, h in enumerate(hmms):
        if opts.cut not in h.cutoffs:
            raise SearchParameterError(f"--cut_{opts.cut.lower()}: {opts.cut} bit thresholds unavailable on model {h.name}")
        c[i] = h.cutoffs[opts.cut]
    opts.cutoffs = c


def frame_suffix(frame: int, nframes: int, translated: bool):
    """Target-name suffix and (strand, start_pos) of frame index f (order of lib/aln.py:241-258)."""
    if translated:
        nfwd = 1 if nframes == 1 else 3
        if nframes == 1 or frame < nfwd:
            return f"_pos{frame + 1}", "+", frame
        return f"_rev_pos{frame - nfwd + 1}", "rev", frame - nfwd
    return ("", "+", None) if frame == 0 else ("_rev", "rev", None)


def _exp_times(lnP: np.ndarray, scale: float, cut: float) -> np.ndarray:
    """exp(lnP) * scale <= cut, decided with math.exp where numpy's exp could differ from it in the last place."""
    v = np.exp(lnP.astype(np.float64)) * scale
    ok = v <= cut
    for i in np.flatnonzero(np.abs(v - cut) <= 1e-9 * abs(cut)):
        ok[i] = math.exp(float(lnP[i])) * scale <= cut
    return ok


def report_index(hits: np.ndarray, target_names, Z: float, opts: SearchOptions, n_hmm: int):
    """Reporting thresholds and HMMER's report order, vectorised: no per-hit Python objects.

    Returns (Zeff, {hmm_index: (idx, domZ)}): idx = indices into `hits` of the reported domains in report order (sequences by
    E-value ascending, ties by name; the domains of a sequence consecutive, in dom_idx order)."""
    Zeff = float(opts.Z) if opts.Z is not None else float(Z)
    out = {}
    if len(hits) == 0:
        return Zeff, out
    order = np.lexsort((hits["dom_idx"], hits["target"], hits["hmm"]))
    hh, tt = hits["hmm"][order], hits["target"][order]
    first = np.r_[True, (hh[1:] != hh[:-1]) | (tt[1:] != tt[:-1])]
    starts = np.flatnonzero(first)
    ends = np.r_[starts[1:], len(order)]
    s_hmm = hh[starts]
    s_lnP = hits["seq_lnP"][order[starts]]
    if opts.cut is not None:
        if opts.cutoffs is None or len(opts.cutoffs) != n_hmm:
            raise SearchParameterError(f"--cut_{opts.cut.lower()}: thresholds not resolved for these profiles (resolve_cutoffs)")
        s_ok = hits["seq_bits"][order[starts]].astype(np.float64) >= opts.cutoffs[s_hmm, 0]
    elif opts.T is not None:
        s_ok = hits["seq_bits"][order[starts]].astype(np.float64) >= opts.T
    else:
        s_ok = _exp_times(s_lnP, Zeff, opts.E)
    d_lnP = hits["dom_lnP"][order]
    d_bits = hits["dom_bits"][order].astype(np.float64)
    # sequences are grouped by profile already (primary sort key)
    hb = np.searchsorted(s_hmm, np.arange(n_hmm + 1))
    for h in range(n_hmm):
        lo, hi = int(hb[h]), int(hb[h + 1])
        if lo == hi:
            continue
        sel = lo + np.flatnonzero(s_ok[lo:hi])
        if len(sel) == 0:
            continue
        domZ = float(opts.domZ) if opts.domZ is not None else float(len(sel))
        k = np.argsort(s_lnP[sel], kind="stable")
        sel = sel[k]
        lp = s_lnP[sel]
        ties = np.flatnonzero(lp[1:] == lp[:-1])
        if len(ties):  # equal E-values: by name (strcmp), as HMMER's sort does
            i = 0
            sel = sel.copy()
            while i < len(sel):
                j = i + 1
                while j < len(sel) and lp[j] == lp[i]:
                    j += 1
                if j - i > 1:
                    grp = sorted(sel[i:j], key=lambda q: target_names(int(tt[starts[q]])).encode())
                    sel[i:j] = grp
                i = j
        # all domains of the selected sequences, sequence by sequence
        cnt = ends[sel] - starts[sel]
        pos = np.repeat(starts[sel] - np.r_[0, np.cumsum(cnt)[:-1]], cnt) + np.arange(int(cnt.sum()))
        if opts.cut is not None:
            d_ok = d_bits[pos] >= opts.cutoffs[h, 1]
        elif opts.domT is not None:
            d_ok = d_bits[pos] >= opts.domT
        else:
            d_ok = _exp_times(d_lnP[pos], domZ, opts.domE)
        out[h] = (order[pos[d_ok]], domZ)
    return Zeff, out


def rows_from_index(hits, idx, domZ, Zeff, model, aligned, target_names, pp=None):
    """Row dicts of one profile (the form write_tblout / write_domtblout / hmmer_text consume) from report_index's indices."""
    rows = []
    ln2 = math.log(2.0)
    names = {}
    for g in idx:
        g = int(g)
        r = hits[g]
        t = int(r["target"])
        if t not in names:
            names[t] = target_names(t)
        dl, sl = float(r["dom_lnP"]), float(r["seq_lnP"])
        rows.append({"hmm": int(r["hmm"]), "target": t, "name": names[t], "dom_idx": int(r["dom_idx"]),
                     "hmmfrom": int(r["hmmfrom"]), "hmmto": int(r["hmmto"]), "alifrom": int(r["iali"]),
                     "alito": int(r["jali"]), "envfrom": int(r["ienv"]), "envto": int(r["jenv"]),
                     "seq_bits": float(r["seq_bits"]), "seq_evalue": math.exp(sl) * Zeff,
                     "seq_bias": float(r["seqbias"]) / ln2,
                     "dom_bits": float(r["dom_bits"]), "dom_cevalue": math.exp(dl) * domZ,
                     "dom_ievalue": math.exp(dl) * Zeff,
                     "dom_bias": float(r["dombias"]) / ln2, "ndom": int(r["ndom_in_seq"]),
                     "oasc": float(r["oasc"]), "domZ": domZ, "Z": Zeff,
                     "model": model[g], "aligned": aligned[g], "pp": pp[g] if pp is not None else None})
    return rows


def report(hits: np.ndarray, model, aligned, target_names, Z: float, opts: SearchOptions, n_hmm: int, pp=None):
    """Apply reporting thresholds and HMMER's report order per profile.

    hits: structured array (capi.HIT_DTYPE) gathered from all shards, ``target`` already global.
    target_names: callable target_index -> full target name (with _pos/_rev suffix).
    Returns {hmm_index: [domain dict, ...]} in report order.
    """
    out = {h: [] for h in range(n_hmm)}
    Zeff, index = report_index(hits, target_names, Z, opts, n_hmm)
    for h, (idx, domZ) in index.items():
        out[h] = rows_from_index(hits, idx, domZ, Zeff, model, aligned, target_names, pp)
    return out


def hit_rows(query: str, rows, ids, nframes: int, translated: bool):
    """hit_info.tsv rows exactly as extract_reads() writes them (lib/aln.py:544-558)."""
    out = []
    for r in rows:
        read, frame = divmod(r["target"], nframes)
        _, strand, start_pos = frame_suffix(frame, nframes, translated)
        out.append((query, ids[read], r["hmmfrom"], r["hmmto"], r["alifrom"], r["alito"], strand, start_pos,
                    r["model"].upper().replace(".", "-"), r["aligned"].upper().replace(".", "-")))
    return out


def write_step_outputs(outdir: str, query: str, rows, ids, raw_seqs, nframes: int, translated: bool, log_lines=()):
    """map/{g}.hit_info.tsv, map/{g}.targets.fas, map/{g}.hmmsearch.log and both ok/ markers
    (lib/aln.py:500-501, 558-561, 569, 663, 673). raw_seqs: callable read_index -> raw nucleotides."""
    os.makedirs(os.path.join(outdir, "map"), exist_ok=True)
    os.makedirs(os.path.join(outdir, "ok"), exist_ok=True)
    seen = set()
    with open(os.path.join(outdir, "map", query + ".hit_info.tsv"), "w") as tsv, \
            open(os.path.join(outdir, "map", query + ".targets.fas"), "w") as fas:
        for r, row in zip(rows, hit_rows(query, rows, ids, nframes, translated)):
            tsv.write("\t".join(str(x) for x in row) + "\n")
            read = r["target"] // nframes
            if row[1] not in seen:
                fas.write(f">{row[1]}\n{raw_seqs(read)}\n")
                seen.add(row[1])
    with open(os.path.join(outdir, "map", query + ".hmmsearch.log"), "w") as log:
        for line in log_lines:
            log.write(line + "\n")
    open(os.path.join(outdir, "ok", f"hmmsearch_{query}.ok"), "w").close()
    open(os.path.join(outdir, "ok", f"extract_reads_{query}.ok"), "w").close()


def write_tblout(path: str, query: str, rows):
    """Minimal HMMER --tblout table for diffing against a real hmmsearch."""
    with open(path, "w") as fh:
        fh.write("#                                                               --- full sequence ---- --- best 1 domain ---- --- domain number estimation ----\n")
        fh.write("# target name        accession  query name           accession    E-value  score  bias   E-value  score  bias   exp reg clu  ov env dom rep inc description of target\n")
        fh.write("#------------------- ---------- -------------------- ---------- --------- ------ ----- --------- ------ -----   --- --- --- --- --- --- --- --- ---------------------\n")
        seen = {}
        for r in rows:
            seen.setdefault(r["target"], []).append(r)
        for doms in seen.values():
            best = max(doms, key=lambda d: d["dom_bits"])
            r = doms[0]
            fh.write(f"{r['name']:<20s} -          {query:<20s} -          {r['seq_evalue']:9.2g} {r['seq_bits']:6.1f} {r['seq_bias']:5.1f} "
                     f"{best['dom_ievalue']:9.2g} {best['dom_bits']:6.1f} {best['dom_bias']:5.1f}   {float(r['ndom']):3.1f} {r['ndom']:3d}   0   0 "
                     f"{r['ndom']:3d} {r['ndom']:3d} {len(doms):3d} {len(doms):3d} -\n")


def write_domtblout(path: str, query: str, qlen: int, rows, target_lengths=None):
    """HMMER --domtblout table: one line per reported domain, 22 columns + description."""
    with open(path, "w") as fh:
        fh.write("#                                                                            --- full sequence --- -------------- this domain -------------   hmm coord   ali coord   env coord\n")
        fh.write("# target name        accession   tlen query name           accession   qlen   E-value  score  bias   #  of  c-Evalue  i-Evalue  score  bias  from    to  from    to  from    to  acc description of target\n")
        fh.write("#------------------- ---------- ----- -------------------- ---------- ----- --------- ------ ----- --- --- --------- --------- ------ ----- ----- ----- ----- ----- ----- ----- ---- ---------------------\n")
        seqs = []
        for r in rows:
            if seqs and seqs[-1][0]["target"] == r["target"]:
                seqs[-1].append(r)
            else:
                seqs.append([r])
        for doms in seqs:
            for n, d in enumerate(doms, 1):
                tlen = target_lengths(d["target"]) if target_lengths else 0
                acc = d.get("oasc", 0.0) / (1.0 + abs(d["envto"] - d["envfrom"]))
                fh.write(f"{d['name']:<20s} -          {tlen:5d} {query:<20s} -          {qlen:5d} {d['seq_evalue']:9.2g} {d['seq_bits']:6.1f} "
                         f"{d['seq_bias']:5.1f} {n:3d} {len(doms):3d} {d['dom_cevalue']:9.2g} {d['dom_ievalue']:9.2g} {d['dom_bits']:6.1f} "
                         f"{d['dom_bias']:5.1f} {d['hmmfrom']:5d} {d['hmmto']:5d} {d['alifrom']:5d} {d['alito']:5d} {d['envfrom']:5d} "
                         f"{d['envto']:5d} {acc:4.2f} -\n")


def shard_bounds(nreads: int, world: int, rank: int):
    """Contiguous read range of one rank (reads shard by chunk; profiles are replicated)."""
    per = (nreads + world - 1) // world
    return min(nreads, rank * per), min(nreads, (rank + 1) * per)


def gather_hits_files(hits: np.ndarray, model, aligned, rank: int, world: int, directory: str, dst: int = 0,
                      timeout: float = 3600.0):
    """Host-side gather of per-rank hit lists for one-process-per-GPU runs (the only cross-rank step; there is no
    collective on the DP path and no torch / NCCL in the product): every rank writes its records into `directory`
    (a path all ranks see, e.g. the PhyloAln output directory), rank `dst` waits for all of them.
    Returns (hits, model, aligned) on rank dst, (None, None, None) elsewhere."""
    import pickle
    import time
    os.makedirs(directory, exist_ok=True)
    tmp = os.path.join(directory, f".hits_rank{rank}.tmp")
    with open(tmp, "wb") as fh:
        pickle.dump((hits.tobytes(), list(model), list(aligned)), fh, protocol=4)
    os.replace(tmp, os.path.join(directory, f"hits_rank{rank}.pkl"))     # atomic: a reader never sees a partial file
    if rank != dst:
        return None, None, None
    parts, t_end = [], time.time() + timeout
    for r in range(world):
        path = os.path.join(directory, f"hits_rank{r}.pkl")
        while not os.path.exists(path):
            if time.time() > t_end:
                raise TimeoutError(f"rank {r} never delivered its hit list ({path})")
            time.sleep(0.01)
        with open(path, "rb") as fh:
            parts.append(pickle.load(fh))
        os.remove(path)
    allhits = np.concatenate([np.frombuffer(b[0], dtype=capi.HIT_DTYPE) for b in parts]) if parts else np.zeros(0, dtype=capi.HIT_DTYPE)
    return allhits, [m for b in parts for m in b[1]], [m for b in parts for m in b[2]]


class ArraySource:
    """Reads already in host memory: (nt uint8, off uint64[n+1])."""

    def __init__(self, nt: np.ndarray, off: np.ndarray):
        self.nt, self.off = nt, off

    def finished(self):
        return True

    def available(self):
        return len(self.off) - 1

    def nbytes(self, a, b):
        return int(self.off[b] - self.off[a])

    def pack(self, a, b):
        sub_nt = self.nt[int(self.off[a]):int(self.off[b])]
        return (sub_nt if len(sub_nt) else np.zeros(1, dtype=np.uint8)), (self.off[a:b + 1] - self.off[a]).astype(np.uint64)


class IngestSource:
    """Records of a (possibly still running) native ingest, in emission order (ingest.Ingest, pa_ingest_*)."""

    def __init__(self, ing):
        self.ing = ing

    def finished(self):
        return self.ing.done()

    def available(self):
        return self.ing.emitted()

    def nbytes(self, a, b):
        return self.ing.emitted_bytes(a, b - a)

    def pack(self, a, b):
        return self.ing.pack_emitted(a, b - a)


CHUNK_BYTE_LIMIT = (1 << 32) - (1 << 20)   # pa_load_reads: 2 * nt + 32 * reads must stay below 4 GiB

_VALID_NT = np.zeros(256, dtype=bool)          # what translate_kernel accepts = Biopython's unambiguous + IUPAC DNA/RNA letters
_VALID_NT[list(b"ACGTURYMKSWHBVDNacgturymkswhbvdn")] = True


def valid_reads_mask(nt: np.ndarray, off: np.ndarray) -> np.ndarray:
    """Per read: does it consist of nucleotide letters only (same test as the device code, csrc/kernels.cuh nt_mask)?"""
    n = len(off) - 1
    bad = ~_VALID_NT[nt[:int(off[-1])]]
    if not bad.any():
        return np.ones(n, dtype=bool)
    cs = np.r_[0, np.cumsum(bad)]
    return (cs[off[1:].astype(np.int64)] - cs[off[:-1].astype(np.int64)]) == 0


class B200Searcher:
    """All profiles x one read set on one or more B200s (one context per GPU, reads sharded by chunk)."""

    def __init__(self, hmms, devices=(0,), options: SearchOptions | None = None, defer: bool = False):
        """defer=True: the per-GPU contexts are created by the first search (so that a running ingest, the parsing of the
        .hmm files and CUDA start-up overlap) instead of here."""
        self.opts = options or SearchOptions()
        self.hmms = [h if isinstance(h, HMM) else read_hmm(h) for h in hmms]
        if self.hmms:
            resolve_cutoffs(self.opts, self.hmms)
        self.devices = list(devices)
        self.ctxs = [None] * len(self.devices)
        self._setup_lock = threading.Lock()
        self._ctx_threads, self._ctx_errors, self._committed = None, [], False
        if not defer:
            self._ensure_contexts()

    def start_contexts(self):
        """Begin creating the CUDA contexts in background threads (CUDA start-up takes ~0.7 s per process and overlaps with
        whatever the caller does next, e.g. parsing the .hmm files); _ensure_contexts() joins them."""
        if getattr(self, "_ctx_threads", None) is not None or all(c is not None for c in self.ctxs):
            return
        self._ctx_errors = []

        def create(k):
            try:
                self.ctxs[k] = capi.Context(self.devices[k])
            except Exception as e:
                self._ctx_errors.append(e)
        self._ctx_threads = [threading.Thread(target=create, args=(k,)) for k in range(len(self.devices))]
        for t in self._ctx_threads:
            t.start()

    def _ensure_contexts(self):
        """One context per GPU (CUDA contexts come up in parallel threads); the profiles are added to the first one and
        committed ONCE for all of them (pa_commit_hmms_multi: configuration and table layout on the host are not repeated
        per GPU, only the upload is)."""
        with self._setup_lock:
            if self._committed:
                return
            import time
            t0 = time.perf_counter()
            self.start_contexts()
            prepared = [capi.prepare_hmm(h) for h in self.hmms]
            for t in self._ctx_threads or []:
                t.join()
            self._ctx_threads = None
            if self._ctx_errors:
                self.close()
                raise self._ctx_errors[0]
            c0 = self.ctxs[0]
            c0.set_opts(F1=self.opts.F1, F2=self.opts.F2, F3=self.opts.F3, do_bias=not self.opts.nobias,
                        do_null2=not self.opts.nonull2, do_max=self.opts.max)
            t1 = time.perf_counter()
            for p in prepared:
                c0.add_hmm(p)
            t2 = time.perf_counter()
            capi.commit_multi(self.ctxs)
            self._committed = True
            self.setup_timing = {"create_contexts_s": t1 - t0, "add_profiles_s": t2 - t1, "commit_and_upload_s": time.perf_counter() - t2}

    def close(self):
        for c in self.ctxs:
            if c is not None:
                c.close()
        self.ctxs = [None] * len(self.ctxs)
        self._committed = False

    def search_reads(self, nt: np.ndarray, off: np.ndarray, **kw):
        """Search concatenated ASCII reads (host arrays). See :meth:`search_source`."""
        return self.search_source(ArraySource(nt, off), **kw)

    def search_source(self, src, mol_type: str = "prot", gencode: int = 1, no_reverse: bool = False, codon: bool = False,
                      chunk_reads: int = 1 << 18, want_pp: bool = False):
        """Search every record of `src` (ArraySource / IngestSource). Returns (hits, model, aligned, Z, nframes, translated);
        with want_pp the PP annotation lines of the alignments are left in ``self.last_pp`` (same order as hits).

        ``mol_type`` follows PhyloAln's -m: 'dna*' searches nucleotide targets, everything else translates
        (lib/aln.py:241-258).  One thread per GPU context pulls the next contiguous range of records from a shared
        cursor as soon as the source has them (a running ingest keeps producing while earlier ranges are searched), at
        most chunk_reads records and 2*nt + 32*reads < 4 GiB per range; ctypes releases the GIL during the calls.  Hit
        target indices are global (record index * nframes + frame).  Z = records * nframes once the source has ended."""
        import time
        translated = not mol_type.startswith("dna")
        results = []
        nframes_box = [1]
        errors = []
        invalid = []
        lock = threading.Lock()
        cursor = [0]
        min_chunk = max(1, chunk_reads // 4)

        def next_range():
            while not errors:
                with lock:
                    fin = src.finished()       # read before available(): once finished the count is final
                    avail = src.available()
                    a = cursor[0]
                    if avail - a >= min_chunk or (fin and avail > a):
                        step = chunk_reads
                        if fin:    # the end is known: finer ranges so that all GPUs finish together
                            step = max(min_chunk, min(chunk_reads, -(-(avail - a) // (2 * len(self.ctxs)))))
                        b = min(avail, a + step)
                        while b - a > 1 and 2 * src.nbytes(a, b) + 32 * (b - a) + 64 >= CHUNK_BYTE_LIMIT:
                            b = a + (b - a) // 2
                        cursor[0] = b
                        return a, b
                    if fin:
                        return None
                time.sleep(0.002)
            return None

        def search_range(ctx, a, b, out):
            """Search records [a, b); on a capacity error (-3) split the range (the library has no fallback).  Records with a
            character Biopython's translate() would reject are left out and remembered (self.last_invalid): whether they
            are an error is the caller's call -- read_fastx's dictionary may drop them before the reference translates."""
            sub_nt, sub_off = src.pack(a, b)
            keep = None
            try:
                try:
                    ctx.load_reads(sub_nt, sub_off, moltype=0 if translated else 1, gencode=gencode, no_reverse=no_reverse,
                                   codon_only=codon)
                except capi.PaError as e:
                    if "(-2)" not in str(e) or "invalid nucleotide" not in str(e):
                        raise
                    valid = valid_reads_mask(sub_nt, sub_off)
                    if valid.all():
                        raise
                    keep = np.flatnonzero(valid)
                    with lock:
                        invalid.extend((a + np.flatnonzero(~valid)).tolist())
                    lens = (sub_off[1:] - sub_off[:-1]).astype(np.int64)[keep]
                    off2 = np.zeros(len(keep) + 1, dtype=np.uint64)
                    off2[1:] = np.cumsum(lens)
                    pos = np.repeat(sub_off[:-1].astype(np.int64)[keep] - off2[:-1].astype(np.int64), lens) + np.arange(int(off2[-1]))
                    nt2 = sub_nt[pos] if len(pos) else np.zeros(1, dtype=np.uint8)
                    ctx.load_reads(nt2, off2, moltype=0 if translated else 1, gencode=gencode, no_reverse=no_reverse, codon_only=codon)
                nf = ctx.num_frames()
                nframes_box[0] = nf
                hits, model, aligned = ctx.search()
            except capi.PaError as e:
                if "(-3)" in str(e) and b - a > 1:
                    mid = (a + b) // 2
                    search_range(ctx, a, mid, out)
                    search_range(ctx, mid, b, out)
                    return
                raise
            pp = ctx.hit_pp(hits) if want_pp else [None] * len(hits)
            hits = hits.copy()
            if keep is None:
                hits["target"] += a * nf
            else:
                hits["target"] = ((a + keep[hits["target"] // nf]) * nf + hits["target"] % nf).astype(hits["target"].dtype)
            out.append((a, hits, model, aligned, pp))

        def work(k):
            try:
                ctx = self.ctxs[k]
                while True:
                    r = next_range()
                    if r is None:
                        break
                    parts = []
                    search_range(ctx, r[0], r[1], parts)
                    with lock:
                        results.extend(parts)
            except Exception as e:  # surfaced after join
                errors.append(e)

        self._ensure_contexts()
        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(self.ctxs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        if cursor[0] == 0 and self.ctxs:      # no records at all: nframes still follows the mode
            nframes_box[0] = ((1 if codon else 3) + (0 if no_reverse else 3)) if translated else (1 if no_reverse else 2)
        nframes = nframes_box[0]
        results.sort(key=lambda r: r[0])
        hits = np.concatenate([r[1] for r in results]) if results else np.zeros(0, dtype=capi.HIT_DTYPE)
        model = [m for r in results for m in r[2]]
        aligned = [m for r in results for m in r[3]]
        self.last_pp = [m for r in results for m in r[4]] if want_pp else None
        nreads = src.available()
        Z = nreads * nframes
        self.last_invalid = sorted(invalid)
        return hits, model, aligned, Z, nframes, translated

    def search_and_report(self, nt, off, ids, **kw):
        hits, model, aligned, Z, nframes, translated = self.search_reads(nt, off, **kw)
        self.last_Z = Z
        if self.last_invalid:
            raise capi.PaError(f"invalid nucleotide character in read {ids[self.last_invalid[0]]} (Biopython would raise TranslationError)")

        def tname(t):
            read, frame = divmod(t, nframes)
            return ids[read] + frame_suffix(frame, nframes, translated)[0]
        rep = report(hits, model, aligned, tname, Z, self.opts, len(self.hmms), pp=self.last_pp)
        return rep, nframes, translated
