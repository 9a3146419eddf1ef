"""HMMER3 text report of a search (`hmmsearch -o`), written from the GPU hit records.

The reference parses this file with Bio.SearchIO 'hmmer3-text' in extract_reads()
(/root/reference/lib/aln.py:511-558); with this emitter and the `hmmsearch` front-end in
phyloaln_b200/bin the UNMODIFIED lib/aln.py can consume GPU results.
Layout follows HMMER 3.3's p7_tophits_Targets / p7_tophits_Domains / p7_alidisplay_Print: per-sequence table,
per-domain table, alignment blocks of the text width (120 columns unless --notextw) with the model line, the
match line (consensus letter on identity, '+' for a positive match log-odds, blank otherwise), the target line
and the posterior-probability line.
"""
from __future__ import annotations

import math

import numpy as np

from .hmmbuild_lite import background
from .hmmio import HMM

BANNER = ("# hmmsearch :: search profile(s) against a sequence database\n"
          "# phyloaln_b200 (HMMER 3.3-compatible report; B200 search path)\n"
          "# - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - -\n")


def _evalue(x: float) -> str:
    return f"{x:9.2g}"


def match_line(hmm: HMM, model: str, aligned: str, hmmfrom: int) -> str:
    """p7_alidisplay's mline: identity -> the model's letter, positive match score -> '+', else blank."""
    bg = background(hmm.abc)
    alpha = hmm.alphabet
    out, k = [], hmmfrom
    for m, a in zip(model, aligned):
        if m == ".":              # insert
            out.append(" ")
            continue
        if a == "-":              # delete
            out.append(" ")
            k += 1
            continue
        if m.upper() == a.upper():
            out.append(m)
        else:
            x = alpha.find(a.upper())
            out.append("+" if x >= 0 and math.exp(-hmm.mat[k, x]) > bg[x] else " ")
        k += 1
    return "".join(out)


def _blocks(hmm: HMM, qname: str, r: dict, textw: int | None):
    model, aligned, pp = r["model"], r["aligned"], r.get("pp")
    mline = match_line(hmm, model, aligned, r["hmmfrom"])
    namew = max(len(qname), len(r["name"]))
    coordw = max(len(str(v)) for v in (r["hmmfrom"], r["hmmto"], r["alifrom"], r["alito"]))
    n = len(model)
    aliw = n if textw is None else max(40, textw - namew - 2 * coordw - 5)
    lines = []
    k, i = r["hmmfrom"], r["alifrom"]
    for pos in range(0, n, aliw):
        mo, al = model[pos:pos + aliw], aligned[pos:pos + aliw]
        nk = sum(1 for c in mo if c != ".")
        ni = sum(1 for c in al if c != "-")
        if hmm.rf:
            rfline = "".join("." if c == "." else "x" for c in mo)
            lines.append(f"  {'':>{namew}} {'':>{coordw}} {rfline} RF")
        if nk:
            lines.append(f"  {qname:>{namew}} {k:>{coordw}d} {mo} {k + nk - 1:<{coordw}d}")
        else:
            lines.append(f"  {qname:>{namew}} {'-':>{coordw}} {mo} {'-':<{coordw}}")
        lines.append(f"  {'':>{namew}} {'':>{coordw}} {mline[pos:pos + aliw]}")
        if ni:
            lines.append(f"  {r['name']:>{namew}} {i:>{coordw}d} {al} {i + ni - 1:<{coordw}d}")
        else:
            lines.append(f"  {r['name']:>{namew}} {'-':>{coordw}} {al} {'-':<{coordw}}")
        if pp:
            lines.append(f"  {'':>{namew}} {'':>{coordw}} {pp[pos:pos + aliw]} PP")
        lines.append("")
        k += nk
        i += ni
    return lines


def write_text_report(fh, hmm: HMM, rows, Z: float, target_lengths=None, params: dict | None = None, incE: float = 0.01,
                      incdomE: float = 0.01, textw: int | None = 120, funnel: dict | None = None, header: dict | None = None):
    """rows: search.report()[h] (report order, domains of a sequence consecutive). fh: text file object."""
    qname = hmm.name
    fh.write(BANNER)
    for k, v in (header or {}).items():
        fh.write(f"# {k + ':':<32s} {v}\n")
    fh.write("# - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - -\n\n")
    fh.write(f"Query:       {qname}  [M={hmm.M}]\n")
    fh.write("Scores for complete sequences (score includes all domains):\n")
    seqs = []
    for r in rows:
        if seqs and seqs[-1][0]["target"] == r["target"]:
            seqs[-1].append(r)
        else:
            seqs.append([r])
    namew = max([8] + [len(d[0]["name"]) for d in seqs])
    fh.write("   --- full sequence ---   --- best 1 domain ---    -#dom-\n")
    fh.write(f"    E-value  score  bias    E-value  score  bias    exp  N  {'Sequence':<{namew}s} Description\n")
    fh.write(f"    ------- ------ -----    ------- ------ -----   ---- --  {'-' * namew} -----------\n")
    if not seqs:
        fh.write("\n   [No hits detected that satisfy reporting thresholds]\n")
    below = False
    for doms in seqs:
        r, best = doms[0], max(doms, key=lambda d: d["dom_bits"])
        if not below and r["seq_evalue"] > incE:
            fh.write("  ------ inclusion threshold ------\n")
            below = True
        fh.write(f"  {_evalue(r['seq_evalue'])} {r['seq_bits']:6.1f} {r['seq_bias']:5.1f}  {_evalue(best['dom_ievalue'])} "
                 f"{best['dom_bits']:6.1f} {best['dom_bias']:5.1f}  {float(r['ndom']):5.1f} {len(doms):2d}  {r['name']:<{namew}s} \n")
    fh.write("\n\nDomain annotation for each sequence (and alignments):\n")
    if not seqs:
        fh.write("\n   [No targets detected that satisfy reporting thresholds]\n")
    for doms in seqs:
        r = doms[0]
        fh.write(f">> {r['name']}  \n")
        fh.write("   #    score  bias  c-Evalue  i-Evalue hmmfrom  hmm to    alifrom  ali to    envfrom  env to     acc\n")
        fh.write(" ---   ------ ----- --------- --------- ------- -------    ------- -------    ------- -------    ----\n")
        L = target_lengths(r["target"]) if target_lengths else None
        for n, d in enumerate(doms, 1):
            inc = "!" if (d["dom_ievalue"] <= incdomE and r["seq_evalue"] <= incE) else "?"
            hb = ("[" if d["hmmfrom"] == 1 else ".") + ("]" if d["hmmto"] == hmm.M else ".")
            ab = ("[" if d["alifrom"] == 1 else ".") + ("]" if L is not None and d["alito"] == L else ".")
            eb = ("[" if d["envfrom"] == 1 else ".") + ("]" if L is not None and d["envto"] == L else ".")
            acc = d["oasc"] / (1.0 + abs(d["envto"] - d["envfrom"]))
            fh.write(f" {n:3d} {inc} {d['dom_bits']:6.1f} {d['dom_bias']:5.1f} {_evalue(d['dom_cevalue'])} {_evalue(d['dom_ievalue'])} "
                     f"{d['hmmfrom']:7d} {d['hmmto']:7d} {hb} {d['alifrom']:7d} {d['alito']:7d} {ab} {d['envfrom']:7d} {d['envto']:7d} {eb} "
                     f"{acc:4.2f}\n")
        fh.write("\n  Alignments for each domain:\n")
        for n, d in enumerate(doms, 1):
            fh.write(f"  == domain {n}  score: {d['dom_bits']:.1f} bits;  conditional E-value: {d['dom_cevalue']:.2g}\n")
            for line in _blocks(hmm, qname, d, textw):
                fh.write(line + "\n")
    fh.write("\n\nInternal pipeline statistics summary:\n-------------------------------------\n")
    fh.write(f"Query model(s):                {1:15d}  ({hmm.M} nodes)\n")
    f = funnel or {}
    fh.write(f"Target sequences:              {int(Z):15d}  ({int(f.get('residues', 0))} residues searched)\n")
    for label, key in (("MSV", "n_past_msv"), ("bias", "n_past_bias"), ("Vit", "n_past_vit"), ("Fwd", "n_past_fwd")):
        if key in f:
            fh.write(f"Passed {label + ' filter:':<24s}{int(f[key]):15d}  ({f[key] / max(Z, 1):.6g})\n")
    domZ = rows[0]["domZ"] if rows else 0
    fh.write(f"Initial search space (Z):      {int(Z):15d}  [actual number of targets]\n")
    fh.write(f"Domain search space  (domZ):   {int(domZ):15d}  [number of targets reported over threshold]\n")
    fh.write("//\n[ok]\n")
