"""Reference-column definition of the hmmer search mode (/root/reference/lib/aln.py:745-792).

After the search the reference writes `map/{g}.ref.fas`: the rows of `ref_hmm/{g}.sto` (the alignment `hmmbuild -O`
annotated) restricted to the columns whose `#=GC RF` character is not '.', i.e. the match columns of the profile --
so reference column c IS profile node c, which is what lets `map_kernel` place a hit by its `hmmfrom` alone -- upper
case, with '.' and '~' written as '-'; and, when the group has a codon alignment, `map/{g}.codon.ref.fas`: the nucleotide
triplets of the same columns taken from the codon alignment after its all-gap columns are removed.  An existing
`ref_hmm/{g}.ref.fas` is copied instead (`:745-749`).

Host code (numpy masks over the byte rows); the file formats are read the way the reference's readers do:
Stockholm as Biopython's `AlignIO.read(..., "stockholm")` (blocks concatenated per id in order of first appearance, '.'
of a row read as '-', every row and `#=GC` line as long as the alignment), aligned FASTA as `lib.read_fastx(..., 'fasta')`
(`lib/library.py:176-198`: id = text up to the first blank, a repeated id keeps its position and restarts its sequence).
"""
from __future__ import annotations

import os

import numpy as np


class StockholmError(ValueError):
    pass


def read_stockholm(path: str):
    """First alignment of a Stockholm file -> (ordered {id: row}, {GC feature: text}).  Raises StockholmError for what
    Biopython's reader rejects (no header, a sequence line without text, rows or `#=GC` lines of unequal length)."""
    rows: dict[str, list] = {}
    gc: dict[str, list] = {}
    with open(path) as fh:
        first = fh.readline()
        if first.strip() != "# STOCKHOLM 1.0":
            raise StockholmError(f"{path}: not a Stockholm 1.0 file")
        ended = False
        for line in fh:
            line = line.strip()
            if line == "//":
                ended = True
                for rest in fh:          # AlignIO.read() accepts exactly one alignment per file
                    if rest.strip():
                        raise StockholmError(f"{path}: more than one alignment in the file")
                break
            if line == "":
                continue
            if line[0] != "#":
                parts = [x.strip() for x in line.split(" ", 1)]
                if len(parts) != 2:
                    raise StockholmError(f"{path}: could not split the line into identifier and sequence: {line[:60]}")
                rows.setdefault(parts[0], []).append(parts[1].replace(".", "-"))
            elif line.startswith("#=GC "):
                f = line[5:].strip().split(None, 1)
                if len(f) == 2:
                    gc.setdefault(f[0], []).append(f[1].strip())
        if not ended and not rows:
            raise StockholmError(f"{path}: no alignment found")
    seqs = {k: "".join(v) for k, v in rows.items()}
    if not seqs:
        raise StockholmError(f"{path}: no sequences")
    n = len(next(iter(seqs.values())))
    for k, v in seqs.items():
        if len(v) != n:
            raise StockholmError(f"{path}: sequences have different lengths ({k})")
    ann = {k: "".join(v) for k, v in gc.items()}
    for k, v in ann.items():
        if len(v) != n:
            raise StockholmError(f"{path}: #=GC {k} has {len(v)} columns, the alignment {n}")
    return seqs, ann


def read_aligned_fasta(path: str) -> dict:
    """`lib.read_fastx(path, 'fasta')` (lib/library.py:176-198) for a plain-text FASTA alignment."""
    seqs: dict[str, str] = {}
    seqid = None
    with open(path) as fh:
        for line in fh:
            line = line.rstrip()
            if line.startswith(">"):
                seqid = line.split(" ")[0].lstrip(">")
                seqs[seqid] = ""
            elif seqid:
                seqs[seqid] += line
    return seqs


_DASH = bytes.maketrans(b".~", b"--")


def strip_all_gap_columns(aln: dict) -> dict:
    """Remove the columns in which every row has '-' (lib/aln.py:755-766); the first row sets the number of columns and
    a shorter row is an error there (IndexError) and here."""
    if not aln:
        return {}
    n = len(next(iter(aln.values())))
    rows = []
    for k, v in aln.items():
        if len(v) < n:
            raise StockholmError(f"codon alignment: row {k} is shorter than the first row")
        rows.append(np.frombuffer(v[:n].encode("latin-1"), dtype=np.uint8))
    keep = np.zeros(n, dtype=bool)
    for r in rows:
        keep |= r != ord("-")
    return {k: r[keep].tobytes().decode("latin-1") for k, r in zip(aln.keys(), rows)}


def reference_columns(sto_path: str, codon_aln_path: str | None = None, rna_to_dna: bool = False):
    """(ref, codon_ref): ordered {id: match-column row} and, with a codon alignment, {id: triplets of the same columns}
    (None without one).  rna_to_dna: the cm* tools' U -> T of lib/aln.py:780-781 (not used by the hmmer mode)."""
    seqs, ann = read_stockholm(sto_path)
    if "RF" not in ann:
        raise StockholmError(f"{sto_path}: no #=GC RF line (hmmbuild -O writes one)")
    keep = np.frombuffer(ann["RF"].encode("latin-1"), dtype=np.uint8) != ord(".")
    cols = np.flatnonzero(keep)
    codon = strip_all_gap_columns(read_aligned_fasta(codon_aln_path)) if codon_aln_path else None
    ref, cref = {}, ({} if codon is not None else None)
    for sid, row in seqs.items():
        b = np.frombuffer(row.encode("latin-1"), dtype=np.uint8)[keep].tobytes().upper().translate(_DASH)
        if rna_to_dna:
            b = b.replace(b"U", b"T")
        ref[sid] = b.decode("latin-1")
        if codon is not None:
            if sid not in codon:
                raise StockholmError(f"codon alignment has no row for {sid}")
            c = codon[sid]
            cref[sid] = "".join(c[3 * i:3 * i + 3] for i in cols.tolist())
    return ref, cref


def reference_columns_step(outdir: str, groups, codon_alns: dict | None = None, rna_to_dna: bool = False) -> dict:
    """Write `map/{g}.ref.fas` (and `map/{g}.codon.ref.fas` for the groups in codon_alns) for every group, as the hmmer
    branch of the reference's `assemble()` does after the search (lib/aln.py:745-792).  Returns {g: number of columns}."""
    os.makedirs(os.path.join(outdir, "map"), exist_ok=True)
    ncols = {}
    for g in groups:
        out = os.path.join(outdir, "map", g + ".ref.fas")
        ready = os.path.join(outdir, "ref_hmm", g + ".ref.fas")
        if os.path.exists(ready):                                  # :745-749, copied line by line
            with open(ready) as src, open(out, "w") as dst:
                for line in src:
                    dst.write(line)
            ncols[g] = None
            continue
        ref, cref = reference_columns(os.path.join(outdir, "ref_hmm", g + ".sto"), (codon_alns or {}).get(g), rna_to_dna)
        with open(out, "w") as fh:
            for sid, s in ref.items():
                fh.write(f">{sid}\n{s}\n")
        if cref is not None:
            with open(os.path.join(outdir, "map", g + ".codon.ref.fas"), "w") as fh:
                for sid, s in cref.items():
                    fh.write(f">{sid}\n{s.upper()}\n")
        ncols[g] = len(next(iter(ref.values()))) if ref else 0
    return ncols
