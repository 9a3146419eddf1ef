"""Command-line front-ends with HMMER's interface, backed by the B200 library (see bin/hmmsearch).

hmmsearch_main() accepts the command line the reference builds at /root/reference/lib/aln.py:664-667
(`hmmsearch -o OUT --tblout TBL --cpu N <hmmsearch_parameters> HMMFILE SEQFILE`) plus the other options of
search.parse_hmmsearch_parameters(); SEQFILE is the already translated `all.target.fasta` (or nucleotide targets
for a DNA profile).  Output: the HMMER3 text report (hmmer_text.write_text_report) on -o / stdout and the
--tblout table.  Errors go to stderr with exit status 1, which the reference turns into "Error in hmmsearch
command" (lib/aln.py:669-672).
"""
from __future__ import annotations

import sys

import numpy as np

from . import capi
from .hmmer_text import write_text_report
from .hmmio import read_hmm
from .search import SearchParameterError, parse_hmmsearch_parameters, report, write_tblout

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
    opts_toks, pos, out, tbl, textw = [], [], None, None, 120
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
    except (SearchParameterError, capi.PaError, OSError, ValueError, ImportError) as e:
        print(f"hmmsearch (phyloaln_b200): {e}", file=sys.stderr)
        return 1
    return 0
