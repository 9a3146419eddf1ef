"""Command-line front-ends with HMMER's interface, backed by the B200 library (see bin/hmmsearch).

hmmsearch_main() accepts the command line the reference builds at /root/reference/lib/aln.py:664-667
(`hmmsearch -o OUT --tblout TBL --cpu N <hmmsearch_parameters> HMMFILE SEQFILE`) plus the other options of
search.parse_hmmsearch_parameters(); SEQFILE is the already translated `all.target.fasta` (or nucleotide targets
for a DNA profile).  Output: the HMMER3 text report (hmmer_text.write_text_report) on -o / stdout and the
--tblout table.  Errors go to stderr with exit status 1, which the reference turns into "Error in hmmsearch
command" (lib/aln.py:669-672).
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import capi
from .hmmer_text import write_text_report
from .hmmio import read_hmm
from .search import SearchParameterError, parse_hmmsearch_parameters, report, resolve_cutoffs, write_domtblout, write_tblout

_WITH_ARG = {"-o", "--tblout", "--domtblout", "--pfamtblout", "-A", "--cpu", "--seed", "--textw", "--tformat", "-E", "--domE", "-T",
             "--domT", "--incE", "--incdomE", "-Z", "--domZ", "--F1", "--F2", "--F3"}


def read_fasta(path: str):
    names, seqs, cur = [], [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if names:
                    seqs.append("".join(cur))
                names.append(line[1:].split()[0] if len(line) > 1 else "")
                cur = []
            elif names:
                cur.append(line.strip())
    if names:
        seqs.append("".join(cur))
    return names, seqs


def hmmsearch_main(argv, device: int = 0) -> int:
    opts_toks, pos, out, tbl, domtbl, textw = [], [], None, None, None, 120
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "-h":
            print(__doc__)
            return 0
        if a in _WITH_ARG:
            if i + 1 >= len(argv):
                print(f"hmmsearch: option {a} needs an argument", file=sys.stderr)
                return 1
            if a == "-o":
                out = argv[i + 1]
            elif a == "--tblout":
                tbl = argv[i + 1]
            elif a == "--domtblout":
                domtbl = argv[i + 1]
            elif a == "--textw":
                textw = int(argv[i + 1])
            opts_toks += [a, argv[i + 1]]
            i += 2
        elif a.startswith("-") and len(a) > 1:
            if a == "--notextw":
                textw = None
            opts_toks.append(a)
            i += 1
        else:
            pos.append(a)
            i += 1
    if len(pos) != 2:
        print("Usage: hmmsearch [options] <hmmfile> <seqdb>", file=sys.stderr)
        return 1
    try:
        opts = parse_hmmsearch_parameters(opts_toks)
        hmm = read_hmm(pos[0])
        resolve_cutoffs(opts, [hmm])
        names, seqs = read_fasta(pos[1])
        ctx = capi.Context(device)
        try:
            ctx.set_opts(F1=opts.F1, F2=opts.F2, F3=opts.F3, do_bias=not opts.nobias, do_null2=not opts.nonull2, do_max=opts.max)
            ctx.add_hmm(hmm)
            ctx.commit()
            n = ctx.load_proteins([s.upper() for s in seqs], abc=0 if hmm.abc == "amino" else 1) if seqs else 0
            if n:
                hits, model, aligned = ctx.search()
                pp = ctx.hit_pp(hits)
                st = ctx.stats()
            else:
                hits, model, aligned, pp, st = np.zeros(0, dtype=capi.HIT_DTYPE), [], [], [], {}
        finally:
            ctx.close()
        rows = report(hits, model, aligned, lambda t: names[t], len(names), opts, 1, pp=pp)[0]
        funnel = dict(st, residues=sum(map(len, seqs)))
        header = {"query HMM file": pos[0], "target sequence database": pos[1]}
        if out:
            header["output directed to file"] = out
        if tbl:
            header["per-seq hits tabular output"] = tbl
        fh = open(out, "w") if out else sys.stdout
        try:
            write_text_report(fh, hmm, rows, len(names), target_lengths=lambda t: len(seqs[t]), incE=opts.incE, incdomE=opts.incdomE,
                              textw=textw, funnel=funnel, header=header)
        finally:
            if out:
                fh.close()
        if tbl:
            write_tblout(tbl, hmm.name, rows)
        if domtbl:
            write_domtblout(domtbl, hmm.name, hmm.M, rows, target_lengths=lambda t: len(seqs[t]))
    except (SearchParameterError, capi.PaError, OSError, ValueError, ImportError) as e:
        print(f"hmmsearch (phyloaln_b200): {e}", file=sys.stderr)
        return 1
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# hmmbuild front-end (bin/hmmbuild): `hmmbuild -O STO -n NAME --cpu N [parameters] HMMFILE MSAFILE`, the command line of
# /root/reference/lib/aln.py:66-77.  NOT a re-implementation of HMMER's model construction: the profile comes from
# hmmbuild_lite (Henikoff position-based weights, symfrac match columns, entropy weighting to HMMER's default target, a
# single-component Dirichlet prior instead of HMMER's mixtures), so its numbers differ from a real hmmbuild's; the
# E-value parameters are calibrated with the GPU filters themselves.  It exists so that a machine without HMMER can run
# the whole default pipeline; with HMMER installed, keep the real hmmbuild.
# ---------------------------------------------------------------------------------------------------------------------
def read_alignment(path: str):
    """Aligned FASTA or single-/multi-block Stockholm -> ordered {id: row}."""
    with open(path) as fh:
        text = fh.read()
    rows: dict[str, str] = {}
    if text.lstrip().startswith("# STOCKHOLM"):
        for line in text.splitlines():
            if not line.strip() or line.startswith("#") or line.startswith("//"):
                continue
            name, seq = line.split()[:2]
            rows[name] = rows.get(name, "") + seq
    else:
        name = None
        for line in text.splitlines():
            if line.startswith(">"):
                name = line[1:].split()[0]
                rows[name] = ""
            elif name is not None:
                rows[name] += line.strip()
    if not rows:
        raise ValueError(f"{path}: no sequences")
    return rows


def guess_alphabet(rows) -> str:
    n = nt = 0
    for s in rows.values():
        for ch in s.upper():
            if ch.isalpha():
                n += 1
                nt += ch in "ACGTUN"
    return "dna" if n and nt / n >= 0.9 else "amino"


def write_stockholm(path: str, name: str, rows: dict, mcols) -> None:
    """hmmbuild -O: match columns upper case with '-' for deletions, insert columns lower case with '.', #=GC RF with an
    'x' per match column -- what lib/aln.py:767-792 reads back (RF != '.' selects the reference columns)."""
    alen = len(next(iter(rows.values())))
    is_match = [False] * alen
    for c in mcols:
        is_match[c] = True
    width = max(len("#=GC RF"), max(len(k) for k in rows)) + 2     # the annotation label must not touch its text
    with open(path, "w") as fh:
        fh.write(f"# STOCKHOLM 1.0\n#=GF ID {name}\n\n")
        for sid, s in rows.items():
            out = []
            for c, ch in enumerate(s):
                gap = not ch.isalpha() and ch != "*"
                if is_match[c]:
                    out.append("-" if gap else ch.upper())
                else:
                    out.append("." if gap else ch.lower())
            fh.write(f"{sid:<{width}s}{''.join(out)}\n")
        fh.write(f"{'#=GC RF':<{width}s}{''.join('x' if m else '.' for m in is_match)}\n//\n")


def gpu_score_fn(device: int = 0):
    """(msv, vit, fwd) raw scores of a profile for a list of sequences, from the filter kernels (E-value calibration)."""
    def score(hmm, seqs):
        ctx = capi.Context(device)
        try:
            ctx.add_hmm(hmm, require_stats=False)
            ctx.commit()
            n = ctx.load_proteins(seqs, abc=0 if hmm.abc == "amino" else 1)
            r = ctx.score_all(0, n, want=("msv", "vit", "fwd"))
            return r["msv"].astype(np.float64), r["vit"].astype(np.float64), r["fwd"].astype(np.float64)
        finally:
            ctx.close()
    return score


def hmmbuild_main(argv, score_fn=None) -> int:
    from . import hmmbuild_lite as hb
    from .hmmio import write_hmm
    with_arg = {"-O", "-n", "-o", "--cpu", "--symfrac", "--informat", "--seed", "--fragthresh", "--wid", "--eset", "--ere", "--esigma",
                "--eid", "--EmL", "--EmN", "--EvL", "--EvN", "--EfL", "--EfN", "--Eft", "--w_beta", "--w_length", "--maxinsertlen"}
    opt, pos, i = {}, [], 0
    while i < len(argv):
        a = argv[i]
        if a in with_arg:
            if i + 1 >= len(argv):
                print(f"hmmbuild: option {a} needs an argument", file=sys.stderr)
                return 1
            opt[a] = argv[i + 1]
            i += 2
        elif a.startswith("-") and len(a) > 1:
            opt[a] = True
            i += 1
        else:
            pos.append(a)
            i += 1
    if len(pos) != 2:
        print("Usage: hmmbuild [-options] <hmmfile_out> <msafile>", file=sys.stderr)
        return 1
    unsupported = [k for k in opt if k in ("--hand", "--fast") and k == "--hand"]
    if unsupported:
        print("hmmbuild (phyloaln_b200): --hand (RF-defined architecture) is not supported", file=sys.stderr)
        return 1
    try:
        rows = read_alignment(pos[1])
        abc = "amino" if "--amino" in opt else "dna" if ("--dna" in opt or "--rna" in opt) else guess_alphabet(rows)
        name = opt.get("-n") or os.path.splitext(os.path.basename(pos[1]))[0]
        hmm, mcols = hb.build_from_alignment(name, rows, abc=abc, symfrac=float(opt.get("--symfrac", 0.5)))
        hb.calibrate(hmm, score_fn or gpu_score_fn())
        write_hmm(hmm, pos[0])
        if "-O" in opt:
            write_stockholm(opt["-O"], name, rows, mcols)
        summary = (f"# hmmbuild (phyloaln_b200 hmmbuild-lite; not HMMER's model construction)\n"
                   f"# idx name                  nseq  alen  mlen eff_nseq\n"
                   f"  1   {name:<20s} {len(rows):5d} {len(next(iter(rows.values()))):5d} {hmm.M:5d} {hmm.effn:8.2f}\n")
        if "-o" in opt:
            with open(opt["-o"], "w") as fh:
                fh.write(summary)
        else:
            sys.stdout.write(summary)
    except (capi.PaError, OSError, ValueError, ImportError) as e:
        print(f"hmmbuild (phyloaln_b200): {e}", file=sys.stderr)
        return 1
    return 0
