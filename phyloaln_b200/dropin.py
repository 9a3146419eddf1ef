"""Drop-in for PhyloAln's per-alignment ``hmmsearch`` + ``extract_reads`` step.

The reference runs, inside ``assemble(g)`` (/root/reference/lib/aln.py:650-674), one ``hmmsearch``
child per alignment over the whole ``all.target.fasta`` and then parses its text report.  The
B200 path inverts the loop nest (SURVEY.md section 8b "Batching reality"): ONE batched search
streams the reads once against ALL profiles, then writes for every alignment g exactly the files
``assemble(g)`` would have produced -- ``map/{g}.hit_info.tsv``, ``map/{g}.targets.fas``,
``map/{g}.hmmsearch.log`` and the markers ``ok/hmmsearch_{g}.ok``, ``ok/extract_reads_{g}.ok`` -- so
the unchanged ``assemble(g)`` finds its resume markers and continues at lib/aln.py:745.
Call :func:`hmmsearch_step` between ``prepare_target`` (lib/main.py:237) and ``assemble_mp`` (:250).

The step is a pipeline: the native ingest (``prepare_target_step(..., wait=False)``) keeps decoding and
de-duplicating the input files while the GPUs already search the first records; only reads with hits ever
become Python strings (ids and sequences are fetched from the ingest's record table for those reads).
"""
from __future__ import annotations

import gzip
import os
import time

import numpy as np

from . import capi
from .hmmer_text import write_text_report
from .hmmio import read_hmm
from .search import (ArraySource, B200Searcher, IngestSource, frame_suffix, parse_hmmsearch_parameters, report_index, resolve_cutoffs,
                     rows_from_index, write_tblout)

MAX_M = 8192   # widest profile of libphyloaln_b200 (csrc/wide.cuh); wider groups are left to the reference's own hmmsearch
PLACEHOLDER = ("# all.target.fasta is not materialised: phyloaln_b200 translates the reads on the GPU (dropin.prepare_target_step).\n"
               "# To run the CPU hmmsearch of the unmodified pipeline on this directory, delete ok/prepare_target.ok and re-run prepare_target.\n")


def read_raw_fasta(path: str):
    """all.raw.fasta (single-line FASTA written by prepare_target, lib/aln.py:443-456).

    Dictionary semantics as in raw2target()/read_fastx (lib/library.py:186-198): a repeated id keeps
    its first position but the LAST sequence wins.
    Returns (ids, nt uint8 array, offsets uint64[n+1])."""
    opener = gzip.open if path.endswith(".gz") else open
    seqs: dict[str, bytes] = {}
    name = None
    with opener(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                name = line[1:].split(b" ")[0].decode()
                seqs[name] = b""
            elif name is not None:
                seqs[name] += line
    ids = list(seqs)
    off = np.zeros(len(ids) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(seqs[i]) for i in ids])
    nt = np.frombuffer(b"".join(seqs[i] for i in ids), dtype=np.uint8)
    return ids, nt, off


def prepare_target_step(outdir: str, rawdata: dict, args, wait: bool = True):
    """Drop-in for ``prepare_target(rawdata, args)`` (lib/main.py:237 -> lib/aln.py:274-457) when the search
    runs on the GPU: the native ingest writes ``all.raw.fasta``, ``total_counts.tsv`` and
    ``merged_redundance.txt`` byte-identically to the reference; translation (raw2target) is not run on the
    host at all -- it is fused into the GPU search -- so ``all.target.fasta`` is a placeholder for the
    reference's final clean-up (lib/aln.py:892-895).  The placeholder is not a sequence file, so an hmmsearch
    run on it by mistake fails loudly instead of reporting zero hits, and ``ok/prepare_target.b200`` marks the
    directory as prepared by this step.  Touches ``ok/prepare_target.ok`` so the unchanged ``prepare_target``
    skips (lib/aln.py:277).  Returns the live :class:`ingest.Ingest` for ``hmmsearch_step``.

    wait=False returns while the input files are still being decoded; ``hmmsearch_step(ingest=...)`` then
    searches the records as they appear and writes the three files and the markers once the ingest has ended."""
    from . import ingest
    os.makedirs(os.path.join(outdir, "ok"), exist_ok=True)
    ing = ingest.prepare_target(rawdata, args, outdir=outdir, wait=wait)
    ing._outdir = outdir
    if wait:
        _finish_prepare_target(outdir, ing)
    return ing


def _finish_prepare_target(outdir: str, ing):
    if getattr(ing, "_files_written", False):
        return
    ing.wait()
    if not os.path.exists(os.path.join(outdir, "total_counts.tsv")) or not os.path.exists(os.path.join(outdir, "all.raw.fasta")):
        ing.write(outdir)
    with open(os.path.join(outdir, "all.target.fasta"), "w") as fh:
        fh.write(PLACEHOLDER)
    open(os.path.join(outdir, "ok", "prepare_target.b200"), "w").close()
    open(os.path.join(outdir, "ok", "prepare_target.ok"), "w").close()
    ing._files_written = True


def reads_from_ingest(ing):
    """(ids, nt, off) with read_raw_fasta's dictionary semantics, straight from the ingest's record table."""
    ids = ing.ids()
    nt, off = ing.pack()
    if len(set(ids)) == len(ids):
        return ids, nt, off
    last = {}
    for k, name in enumerate(ids):
        last[name] = k                     # the LAST record of an id wins, at the id's first position
    keep = list(last)                      # dict order = first appearance
    raw = nt.tobytes()
    seqs = [raw[int(off[last[n]]):int(off[last[n] + 1])] for n in keep]
    off2 = np.zeros(len(keep) + 1, dtype=np.uint64)
    off2[1:] = np.cumsum([len(x) for x in seqs])
    return keep, np.frombuffer(b"".join(seqs) or b"\0", dtype=np.uint8), off2


def _dictionary_view(ing, hits, nframes):
    """read_fastx's dictionary over all.raw.fasta (lib/library.py:186-198) applied AFTER the search of every emitted
    record: one target set per distinct id, at the id's first position in the file, with the sequence of its LAST record.
    Returns (hits of the surviving records with `target` renumbered to dictionary positions, number of entries,
    record index of every entry)."""
    n = ing.n
    order = ing.final_order()
    order = np.arange(n, dtype=np.uint64) if order is None else order      # final position -> emission index
    ids = ing.ids_of(order)
    last = {}
    for pos, name in enumerate(ids):
        last[name] = pos
    entry_rec = np.array([int(order[last[name]]) for name in last], dtype=np.int64)   # dict order = first appearance
    entry_of_rec = np.full(n, -1, dtype=np.int64)
    entry_of_rec[entry_rec] = np.arange(len(entry_rec))
    rec = hits["target"] // nframes
    keep = entry_of_rec[rec] >= 0
    out = hits[keep].copy()
    out["target"] = (entry_of_rec[rec[keep]] * nframes + (hits["target"][keep] % nframes)).astype(hits["target"].dtype)
    return out, np.flatnonzero(keep), len(entry_rec), entry_rec


def _pa(s: str) -> str:
    return s.upper().replace(".", "-")    # extract_reads: str(...).upper().replace('.', '-') (lib/aln.py:531-532)


def hmmsearch_step(outdir: str, groups, mol_type: str = "prot", gencode: int = 1, no_reverse: bool = False,
                   codon: bool = False, hmmsearch_parameters=(), devices=(0,), raw_fasta: str | None = None,
                   write_tbl: bool = True, ingest=None, write_text: bool = False, chunk_reads: int = 1 << 18):
    """Batched replacement of ``for g in groups: hmmsearch g.hmm all.target.fasta; extract_reads(g)``.

    ingest: the object returned by :func:`prepare_target_step` (may still be running: the search overlaps it); without it
    the reads come from ``raw_fasta`` / ``{outdir}/all.raw.fasta``.  write_text additionally writes map/{g}.hmmsearch.txt,
    the HMMER3 text report of the reference's `-o` (lib/aln.py:664), from the same hit records.  Groups whose profile is
    wider than this library's limit are skipped (no ok markers), so the unmodified ``assemble(g)`` runs the reference's
    own hmmsearch for them.  Returns {g: number of hit_info rows}.  Raises on any failure (the reference returns (1, g) per
    failed group, lib/aln.py:669-672; here the whole batch fails loudly -- there is no CPU fallback)."""
    groups = list(groups)
    opts = parse_hmmsearch_parameters(hmmsearch_parameters)
    todo = [g for g in groups if not (os.path.isfile(os.path.join(outdir, "ok", f"hmmsearch_{g}.ok")) and
                                      os.path.isfile(os.path.join(outdir, "ok", f"extract_reads_{g}.ok")))]
    counts = {}
    tm = {}
    hmmsearch_step.last_timing = tm
    hmmsearch_step.skipped = []
    if not todo:
        if ingest is not None:
            _finish_prepare_target(outdir, ingest)
        return counts
    t0 = time.perf_counter()
    searcher = B200Searcher([], devices=devices, options=opts, defer=True)
    searcher.start_contexts()          # CUDA start-up on every GPU runs while the profiles are parsed
    hmms, kept = [], []
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:    # the native reader releases the GIL
        parsed = list(ex.map(lambda g: read_hmm(os.path.join(outdir, "ref_hmm", g + ".hmm")), todo))
    for g, h in zip(todo, parsed):
        if h.M > MAX_M:
            hmmsearch_step.skipped.append(g)
            continue
        hmms.append(h)
        kept.append(g)
    todo = kept
    try:
        resolve_cutoffs(opts, hmms)    # --cut_ga / --cut_tc / --cut_nc: every profile's own thresholds, or an error
    except Exception:
        for t in searcher._ctx_threads or []:      # the contexts being created beside the parsing are not left behind
            t.join()
        searcher.close()
        raise
    tm["read_hmms_s"] = time.perf_counter() - t0
    # '--codon' forces no_reverse in the reference (lib/main.py:196-198)
    if codon:
        no_reverse = True
    if not todo:
        for t in searcher._ctx_threads or []:
            t.join()
        searcher.close()
        if ingest is not None:
            _finish_prepare_target(outdir, ingest)
        return counts
    t0 = time.perf_counter()
    ids = nt = off = None
    fin_thread, fin_err = None, []
    if ingest is not None:
        src = IngestSource(ingest)

        def finish():     # all.raw.fasta & co. are written as soon as the ingest has ended, next to the last searches
            try:
                _finish_prepare_target(outdir, ingest)
            except Exception as e:
                fin_err.append(e)
        import threading
        fin_thread = threading.Thread(target=finish)
        fin_thread.start()
    else:
        raw_fasta = raw_fasta or os.path.join(outdir, "all.raw.fasta")
        from .ingest import load_raw_fasta     # native loader; read_raw_fasta above is its plain-Python statement
        ids, nt, off = load_raw_fasta(raw_fasta)
        src = ArraySource(nt, off)
    tm["load_reads_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    searcher.hmms = hmms               # configured, laid out and uploaded once for all GPUs when the search starts
    tm["upload_profiles_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    try:
        hits, model, aligned, Z, nframes, translated = searcher.search_source(
            src, mol_type=mol_type, gencode=gencode, no_reverse=no_reverse, codon=codon, chunk_reads=chunk_reads, want_pp=write_text)
        pp = searcher.last_pp
        invalid = searcher.last_invalid
        stats = [c.stats() for c in searcher.ctxs if c is not None]
        tm["setup"] = getattr(searcher, "setup_timing", None)
    finally:
        searcher.close()
    tm["search_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    nrec = src.available()
    sel = None
    if ingest is not None:
        ingest.wait()
        tm["finish_ingest_s"] = time.perf_counter() - t0
        if ingest.repeated_ids() > 0:    # read_fastx keeps one entry per id (config 1 does this: SURVEY B.5)
            hits, sel, nent, entry_rec = _dictionary_view(ingest, hits, nframes)
            Z = nent * nframes
            nrec = nent
        else:
            entry_rec = None
    else:
        entry_rec = None
    if invalid:    # a record with a non-nucleotide letter is an error only if the reference would translate it
        live = set(entry_rec.tolist()) if (ingest is not None and entry_rec is not None) else None
        bad = [r for r in invalid if live is None or r in live]
        if bad:
            name = ingest.ids_of(bad[:1])[0] if ingest is not None else ids[bad[0]]
            raise capi.PaError(f"invalid nucleotide character in read {name} (Biopython would raise TranslationError)")
    if sel is not None:
        model = [model[i] for i in sel]
        aligned = [aligned[i] for i in sel]
        pp = [pp[i] for i in sel] if pp is not None else None
    # ids and raw sequences of the reads that have a hit -- the only ones that ever become Python strings
    t0 = time.perf_counter()
    hit_reads = np.unique(hits["target"] // nframes) if len(hits) else np.zeros(0, dtype=np.int64)
    if ingest is not None:
        recs = hit_reads if entry_rec is None else entry_rec[hit_reads]
        id_of = dict(zip(hit_reads.tolist(), ingest.ids_of(recs)))
        seq_of = dict(zip(hit_reads.tolist(), ingest.seqs_of(recs)))
    else:
        id_of = {int(r): ids[int(r)] for r in hit_reads}
        seq_of = {int(r): nt[int(off[r]):int(off[r + 1])].tobytes().decode() for r in hit_reads}     # no copy of the whole read set
    suffixes = [frame_suffix(f, nframes, translated) for f in range(nframes)]

    def tname(t):
        read, frame = divmod(int(t), nframes)
        return id_of[read] + suffixes[frame][0]
    Zeff, index = report_index(hits, tname, Z, opts, len(hmms))
    tm["report_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    os.makedirs(os.path.join(outdir, "map"), exist_ok=True)
    os.makedirs(os.path.join(outdir, "ok"), exist_ok=True)
    empty = np.zeros(0, dtype=np.int64)
    for hi, g in enumerate(todo):
        idx, domZ = index.get(hi, (empty, 0.0))
        tgt = hits["target"][idx]
        reads, frames = tgt // nframes, tgt % nframes
        cols = [hits[f][idx].tolist() for f in ("hmmfrom", "hmmto", "iali", "jali")]
        tsv, fas, seen = [], [], set()
        for k, (gi, read, fr) in enumerate(zip(idx.tolist(), reads.tolist(), frames.tolist())):
            rid = id_of[read]
            _, strand, start_pos = suffixes[fr]
            tsv.append(f"{g}\t{rid}\t{cols[0][k]}\t{cols[1][k]}\t{cols[2][k]}\t{cols[3][k]}\t{strand}\t{start_pos}\t{_pa(model[gi])}\t{_pa(aligned[gi])}\n")
            if rid not in seen:      # first time a read is seen: its raw nucleotides (lib/aln.py:559-561)
                seen.add(rid)
                fas.append(f">{rid}\n{seq_of[read]}\n")
        with open(os.path.join(outdir, "map", g + ".hit_info.tsv"), "w") as fh:
            fh.write("".join(tsv))
        with open(os.path.join(outdir, "map", g + ".targets.fas"), "w") as fh:
            fh.write("".join(fas))
        with open(os.path.join(outdir, "map", g + ".hmmsearch.log"), "w") as fh:
            fh.write(f"phyloaln_b200 batched search: {nrec} reads x {len(todo)} profiles on GPU(s) {list(devices)}\n"
                     f"Z = {Z} targets; reported domains for {g}: {len(idx)}\nstats: {stats}\n")
        if write_tbl or write_text:
            rows = rows_from_index(hits, idx, domZ, Zeff, model, aligned, tname, pp)
            if write_tbl:
                write_tblout(os.path.join(outdir, "map", g + ".hmmsearch.tbl"), g, rows)
            if write_text:
                with open(os.path.join(outdir, "map", g + ".hmmsearch.txt"), "w") as fh:
                    write_text_report(fh, hmms[hi], rows, Z, incE=opts.incE, incdomE=opts.incdomE,
                                      header={"query HMM file": os.path.join("ref_hmm", g + ".hmm"), "target sequence database": "all.target.fasta"})
        open(os.path.join(outdir, "ok", f"hmmsearch_{g}.ok"), "w").close()
        open(os.path.join(outdir, "ok", f"extract_reads_{g}.ok"), "w").close()
        counts[g] = len(idx)
    tm["write_outputs_s"] = time.perf_counter() - t0
    if fin_thread is not None:
        t0 = time.perf_counter()
        fin_thread.join()
        tm["wait_raw_fasta_written_s"] = time.perf_counter() - t0
        if fin_err:
            raise fin_err[0]
    hmmsearch_step.last_stats = stats
    return counts
