"""Drop-in for PhyloAln's per-alignment ``hmmsearch`` + ``extract_reads`` step.

The reference runs, inside ``assemble(g)`` (/root/reference/lib/aln.py:650-674), one ``hmmsearch``
child per alignment over the whole ``all.target.fasta`` and then parses its text report.  The
B200 path inverts the loop nest (SURVEY.md section 8b "Batching reality"): ONE batched search
streams the reads once against ALL profiles, then writes for every alignment g exactly the files
``assemble(g)`` would have produced -- ``map/{g}.hit_info.tsv``, ``map/{g}.targets.fas``,
``map/{g}.hmmsearch.log`` and the markers ``ok/hmmsearch_{g}.ok``, ``ok/extract_reads_{g}.ok`` -- so
the unchanged ``assemble(g)`` finds its resume markers and continues at lib/aln.py:745.
Call :func:`hmmsearch_step` between ``prepare_target`` (lib/main.py:237) and ``assemble_mp`` (:250).
"""
from __future__ import annotations

import gzip
import os

import numpy as np

from .hmmer_text import write_text_report
from .hmmio import read_hmm
from .search import B200Searcher, parse_hmmsearch_parameters, write_step_outputs, write_tblout


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


def prepare_target_step(outdir: str, rawdata: dict, args):
    """Drop-in for ``prepare_target(rawdata, args)`` (lib/main.py:237 -> lib/aln.py:274-457) when the search
    runs on the GPU: the native ingest writes ``all.raw.fasta``, ``total_counts.tsv`` and
    ``merged_redundance.txt`` byte-identically to the reference; translation (raw2target) is not run on the
    host at all -- it is fused into the GPU search -- so ``all.target.fasta`` is left as an empty placeholder
    for the reference's final clean-up (lib/aln.py:892-895).  Touches ``ok/prepare_target.ok`` so the unchanged
    ``prepare_target`` skips (lib/aln.py:277).  Returns the live :class:`ingest.Ingest` for ``hmmsearch_step``."""
    from . import ingest
    os.makedirs(os.path.join(outdir, "ok"), exist_ok=True)
    ing = ingest.prepare_target(rawdata, args, outdir=outdir)
    open(os.path.join(outdir, "all.target.fasta"), "w").close()
    open(os.path.join(outdir, "ok", "prepare_target.ok"), "w").close()
    return ing


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


def hmmsearch_step(outdir: str, groups, mol_type: str = "prot", gencode: int = 1, no_reverse: bool = False,
                   codon: bool = False, hmmsearch_parameters=(), devices=(0,), raw_fasta: str | None = None,
                   write_tbl: bool = True, ingest=None, write_text: bool = False):
    """Batched replacement of ``for g in groups: hmmsearch g.hmm all.target.fasta; extract_reads(g)``.

    write_text additionally writes map/{g}.hmmsearch.txt, the HMMER3 text report of the reference's `-o`
    (lib/aln.py:664), from the same hit records.  Returns {g: number of hit_info rows}.  Raises on any failure (the reference returns (1, g) per
    failed group, lib/aln.py:669-672; here the whole batch fails loudly -- there is no CPU fallback)."""
    groups = list(groups)
    opts = parse_hmmsearch_parameters(hmmsearch_parameters)
    todo = [g for g in groups if not (os.path.isfile(os.path.join(outdir, "ok", f"hmmsearch_{g}.ok")) and
                                      os.path.isfile(os.path.join(outdir, "ok", f"extract_reads_{g}.ok")))]
    counts = {}
    if not todo:
        return counts
    import time
    tm = {}
    t0 = time.perf_counter()
    hmms = [read_hmm(os.path.join(outdir, "ref_hmm", g + ".hmm")) for g in todo]
    tm["read_hmms_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    if ingest is not None:
        ids, nt, off = reads_from_ingest(ingest)
    else:
        raw_fasta = raw_fasta or os.path.join(outdir, "all.raw.fasta")
        from .ingest import load_raw_fasta     # native loader; read_raw_fasta below is its plain-Python statement
        ids, nt, off = load_raw_fasta(raw_fasta)
    # '--codon' forces no_reverse in the reference (lib/main.py:196-198)
    if codon:
        no_reverse = True
    tm["load_reads_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    searcher = B200Searcher(hmms, devices=devices, options=opts)
    tm["upload_profiles_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    try:
        rep, nframes, translated = searcher.search_and_report(nt, off, ids, mol_type=mol_type, gencode=gencode,
                                                              no_reverse=no_reverse, codon=codon, want_pp=write_text)
        stats = [c.stats() for c in searcher.ctxs]
    finally:
        searcher.close()
    tm["search_and_report_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    raw = nt.tobytes()

    def raw_seq(read):
        return raw[int(off[read]):int(off[read + 1])].decode()
    for hi, g in enumerate(todo):
        rows = rep[hi]
        log = [f"phyloaln_b200 batched search: {len(ids)} reads x {len(todo)} profiles on GPU(s) {list(devices)}",
               f"Z = {len(ids) * nframes} targets; reported domains for {g}: {len(rows)}", f"stats: {stats}"]
        write_step_outputs(outdir, g, rows, ids, raw_seq, nframes, translated, log)
        if write_tbl:
            write_tblout(os.path.join(outdir, "map", g + ".hmmsearch.tbl"), g, rows)
        if write_text:
            with open(os.path.join(outdir, "map", g + ".hmmsearch.txt"), "w") as fh:
                write_text_report(fh, hmms[hi], rows, searcher.last_Z, incE=opts.incE, incdomE=opts.incdomE,
                                  header={"query HMM file": os.path.join("ref_hmm", g + ".hmm"), "target sequence database": "all.target.fasta"})
        counts[g] = len(rows)
    tm["write_outputs_s"] = time.perf_counter() - t0
    hmmsearch_step.last_timing = tm
    return counts
