#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run in the BUILD container only:
needs /root/reference, which does not exist on the GPU box).

What is pinned by the REFERENCE's own Python (imported from /root/reference with the test-only Bio
shim in tests/bio_shim):
  config1/all.raw.fasta.gz       prepare_target() on the four tests/*.fq.gz (tagging, dedup)  [lib/aln.py:274-457]
  config1/target_head.fasta      first records of all.target.fasta from raw2target()          [lib/aln.py:233-259]
  config1/target.sha256          checksum of the whole all.target.fasta (dictionary semantics)
  config1/total_counts.tsv, merged_redundance.txt
  mapping_golden.json            pretreat_hits() + read_hit.__init__() on hit rows             [lib/assemble.py:767-873, 13-306]
What is pinned by the ORACLE (the HMMER boundary has no golden vectors anywhere, "parity unpinned"):
  config1/ref_hmm/*.hmm          profiles from the tests/ref CDS sets (hmmbuild-lite + oracle calibration)
  config1/ref/*.ref.fas          match-column reference alignments (protein) + codon alignments
  config1/hit_info/*.tsv         oracle pipeline + report rows for config 1
"""
import gzip
import hashlib
import json
import os
import shutil
import sys
import tempfile
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "bio_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "lib"))

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402
from oracle import report as orep  # noqa: E402
from phyloaln_b200 import hmmbuild_lite as hb  # noqa: E402
from phyloaln_b200 import hmmio  # noqa: E402

OUT = os.path.join(HERE, "config1")


def read_fasta(path):
    seqs, name = {}, None
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs[name] = ""
        elif name:
            seqs[name] += line
    return seqs


def nw_align(a, b, match=2, mismatch=-1, gap=-3):
    """Global alignment of b onto a; returns gapped (a, b)."""
    n, m = len(a), len(b)
    S = np.zeros((n + 1, m + 1), dtype=np.int32)
    S[:, 0] = gap * np.arange(n + 1)
    S[0, :] = gap * np.arange(m + 1)
    for i in range(1, n + 1):
        ai = a[i - 1]
        for j in range(1, m + 1):
            S[i, j] = max(S[i - 1, j - 1] + (match if ai == b[j - 1] else mismatch), S[i - 1, j] + gap, S[i, j - 1] + gap)
    ga, gb = [], []
    i, j = n, m
    while i > 0 or j > 0:
        if i > 0 and j > 0 and S[i, j] == S[i - 1, j - 1] + (match if a[i - 1] == b[j - 1] else mismatch):
            ga.append(a[i - 1]); gb.append(b[j - 1]); i -= 1; j -= 1
        elif i > 0 and S[i, j] == S[i - 1, j] + gap:
            ga.append(a[i - 1]); gb.append("-"); i -= 1
        else:
            ga.append("-"); gb.append(b[j - 1]); j -= 1
    return "".join(reversed(ga)), "".join(reversed(gb))


def center_star(prots):
    """Tiny stand-in for MAFFT (tests/run_test.sh:8 uses alignseq.pl -> linsi): center-star MSA."""
    names = list(prots)
    center = max(names, key=lambda n: len(prots[n]))
    c = prots[center]
    pair = {n: nw_align(c, prots[n]) for n in names if n != center}
    # number of insertion columns needed after each center position (0 = before the first)
    ins = [0] * (len(c) + 1)
    for ga, gb in pair.values():
        pos, run = 0, 0
        for x in ga:
            if x == "-":
                run += 1
            else:
                ins[pos] = max(ins[pos], run)
                run = 0
                pos += 1
        ins[pos] = max(ins[pos], run)
    out = {}
    for n in names:
        if n == center:
            ga, gb = c, c
        else:
            ga, gb = pair[n]
        row, pos, k = [], 0, 0
        buf = []
        for x, y in zip(ga, gb):
            if x == "-":
                buf.append(y)
            else:
                row.append("".join(buf).ljust(ins[pos], "-"))
                buf = []
                row.append(y)
                pos += 1
        row.append("".join(buf).ljust(ins[pos], "-"))
        out[n] = "".join(row)
    assert len({len(v) for v in out.values()}) == 1
    return out


def build_profiles():
    os.makedirs(os.path.join(OUT, "ref_hmm"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "ref"), exist_ok=True)
    profiles = {}
    for fn in sorted(os.listdir(os.path.join(REF, "tests", "ref"))):
        g = fn.split(".")[0]
        cds = {k: v.upper() for k, v in read_fasta(os.path.join(REF, "tests", "ref", fn)).items()}
        prots = {k: O.translate(v, 1).rstrip("*") for k, v in cds.items()}
        aln = center_star(prots)
        hmm, mcols = hb.build_from_alignment(g, aln, "amino")
        P = O.Profile.from_hmm(hmm)
        hb.calibrate(hmm, lambda hm, seqs, P=P: P.score_all(seqs), seed=42, n=200)
        hmmio.write_hmm(hmm, os.path.join(OUT, "ref_hmm", g + ".hmm"))
        # reference-column definition (lib/aln.py:750-792): match columns only; codon triplets of the same columns
        with open(os.path.join(OUT, "ref", g + ".ref.fas"), "w") as fa, open(os.path.join(OUT, "ref", g + ".codon.ref.fas"), "w") as fc:
            for name, row in aln.items():
                fa.write(f">{name}\n{''.join(row[c] for c in mcols)}\n")
                cod, p = [], 0
                for c in range(len(row)):
                    if row[c] == "-":
                        tri = "---"
                    else:
                        tri = cds[name][3 * p:3 * p + 3]
                        p += 1
                    if c in set(mcols):
                        cod.append(tri)
                fc.write(f">{name}\n{''.join(cod)}\n")
        profiles[g] = hmmio.read_hmm(os.path.join(OUT, "ref_hmm", g + ".hmm"))  # exactly what the tests will load
        print("profile", g, "M =", hmm.M, hmm.stats)
    return profiles


def run_reference_prepare_target(tmp):
    import lib.aln as al  # the reference's own code
    args = Namespace(file_format="guess", split_len=None, split_slide=None, merge_len=250, mol_type="codon", gencode=1,
                     no_reverse=False, codon=False, cpu=1)
    rawdata = {"DMELA": [os.path.join(REF, "tests", "DMELA_1.fq.gz"), os.path.join(REF, "tests", "DMELA_2.fq.gz")],
               "DWILL": [os.path.join(REF, "tests", "DWILL_1.fq.gz"), os.path.join(REF, "tests", "DWILL_2.fq.gz")]}
    cwd = os.getcwd()
    os.chdir(tmp)
    os.makedirs("ok", exist_ok=True)
    try:
        al.prepare_target(rawdata, args)
    finally:
        os.chdir(cwd)
    with open(os.path.join(tmp, "all.raw.fasta"), "rb") as f, gzip.GzipFile(os.path.join(OUT, "all.raw.fasta.gz"), "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    tgt = open(os.path.join(tmp, "all.target.fasta")).read()
    open(os.path.join(OUT, "target.sha256"), "w").write(hashlib.sha256(tgt.encode()).hexdigest() + "\n")
    open(os.path.join(OUT, "target_head.fasta"), "w").write("".join(tgt.splitlines(True)[:240]))
    for f in ("total_counts.tsv", "merged_redundance.txt"):
        shutil.copy(os.path.join(tmp, f), os.path.join(OUT, f))
    names, seqs = [], []
    for line in tgt.splitlines():
        if line.startswith(">"):
            names.append(line[1:])
        else:
            seqs.append(line)
    return names, seqs


def oracle_hit_info(profiles, names, seqs):
    os.makedirs(os.path.join(OUT, "hit_info"), exist_ok=True)
    Z = len(names)
    allrows = {}
    for g, hmm in profiles.items():
        P = O.Profile.from_hmm(hmm)
        hits, _ = P.search(seqs, nthreads=os.cpu_count() or 1, want_trace=False)
        rep = orep.apply_report(hits, names, Z)
        rows = orep.rows_from_hits(g, rep, names)
        with open(os.path.join(OUT, "hit_info", g + ".hit_info.tsv"), "w") as fh:
            for r in rows:
                fh.write("\t".join(str(x) for x in r) + "\n")
        with open(os.path.join(OUT, "hit_info", g + ".scores.tsv"), "w") as fh:  # scores/E-values for tolerance checks
            for h in rep:
                fh.write(f"{names[h['target']]}\t{h['dom_idx']}\t{h['seq_bits']:.4f}\t{h['seq_lnP']:.6f}\t{h['dom_bits']:.4f}\t{h['dom_lnP']:.6f}\n")
        allrows[g] = rows
        print("hit_info", g, len(hits), "domains,", len(rows), "reported rows")
    return allrows


def mapping_golden(allrows):
    """Reference pretreat_hits + read_hit on the golden rows (codon mode, -u SLEBA, defaults) plus KATs."""
    import assemble as asm  # /root/reference/lib/assemble.py
    raw = {}
    name = None
    for line in gzip.open(os.path.join(OUT, "all.raw.fasta.gz"), "rt"):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:]
        else:
            raw[name] = line
    cases = []
    for g, rows in allrows.items():
        refseqs = read_fasta(os.path.join(OUT, "ref", g + ".ref.fas"))
        site_comp, out_score = asm.get_ref_score(refseqs, outgroup=["SLEBA"], sep=".", unknow="X")
        comp = {str(c): v for c, v in site_comp.items()}
        for r in rows[:120]:
            info = [str(x) for x in r]
            new, st, et = asm.pretreat_hits(list(info), site_comp, out_score, no_assemble=False, trim_pos=5, pos_freq=0.2)
            case = {"group": g, "row": info, "read": raw[r[1]]}
            if new is None:
                case["keep"] = False
            else:
                hit = asm.read_hit([str(x) for x in new[2:]], [r[1]], raw[r[1]].upper(), site_comp, st, et, 0, 0.2, 1, None)
                cols = range(hit.qstart, hit.qend + 1)
                case.update(keep=True, start_trim=st, end_trim=et, trimmed=[str(x) for x in new[2:6]],
                            qstart=hit.qstart, qend=hit.qend, base="".join(hit.base[c] for c in cols),
                            codon="".join("".join(list(hit.seq_comp[c][i].keys())[0] for i in (1, 2, 3)) for c in cols),
                            seqids=hit.seqids[r[1]], unmap=sorted(hit.unmap[r[1]].keys()))
            cases.append(case)
        cases.append({"group": g, "site_composition": comp})
    # nucleotide-mode KAT-5 of SURVEY B.4
    site = {i + 1: {c: 1} for i, c in enumerate("ATGAAAACC")}
    row = ["g", "rn", "1", "9", "3", "12", "rev", "None", "ATGAAA-ACC", "ATGAAAGACC"]
    new, st, et = asm.pretreat_hits(list(row), site, {}, trim_pos=5, pos_freq=0.2)
    hit = asm.read_hit([str(x) for x in new[2:]], ["rn"], "AGGTCTTTCATGG", site, st, et, 0, 0.2, 1, None)
    cases.append({"group": "KAT5", "row": row, "read": "AGGTCTTTCATGG", "keep": True, "start_trim": st, "end_trim": et,
                  "trimmed": [str(x) for x in new[2:6]], "qstart": hit.qstart, "qend": hit.qend,
                  "base": "".join(hit.base[c] for c in range(hit.qstart, hit.qend + 1)), "codon": None, "seqids": hit.seqids["rn"],
                  "unmap": sorted(hit.unmap["rn"].keys()), "site_composition_inline": {str(k): v for k, v in site.items()}})
    # synthetic stress rows through the reference code: indels, low-frequency ends, both strands, all frames
    rng = np.random.default_rng(11)
    for g, rows in allrows.items():
        refseqs = read_fasta(os.path.join(OUT, "ref", g + ".ref.fas"))
        site_comp, out_score = asm.get_ref_score(refseqs, outgroup=["SLEBA"], sep=".", unknow="X")
        M = len(next(iter(refseqs.values())))
        cons = "".join(max(site_comp[c + 1], key=site_comp[c + 1].get) if site_comp[c + 1] else "A" for c in range(M))
        for t in range(40):
            span = int(rng.integers(4, 30))
            a = int(rng.integers(1, M - span))
            model, ali, prot = [], [], []
            for c in range(a, a + span):
                u = rng.random()
                aa = cons[c - 1] if rng.random() > 0.25 else "ACDEFGHIKLMNPQRSTVWY"[int(rng.integers(20))]
                if aa in "-X":
                    aa = "A"
                if u < 0.08 and 0 < c - a < span - 1:
                    model.append(cons[c - 1] if cons[c - 1] != "-" else "A"); ali.append("-")
                else:
                    model.append(cons[c - 1] if cons[c - 1] != "-" else "A"); ali.append(aa); prot.append(aa)
                    if u > 0.92 and 0 < c - a < span - 1:
                        model.append("-"); ali.append("W"); prot.append("W")
            fr = int(rng.integers(0, 3))
            from phyloaln_b200.synth import back_translate
            nt = "ACGT"[int(rng.integers(4))] * fr + back_translate("".join(prot)) + "ACG"[:int(rng.integers(0, 3))]
            strand = "+" if rng.random() < 0.5 else "rev"
            read = nt if strand == "+" else O.revcomp(nt)
            row = [g, f"syn{t}", str(a), str(a + span - 1), "1", str(len(prot)), strand, str(fr), "".join(model), "".join(ali)]
            new, st, et = asm.pretreat_hits(list(row), site_comp, out_score, no_assemble=False, trim_pos=5, pos_freq=0.2)
            case = {"group": g, "row": row, "read": read}
            if new is None:
                case["keep"] = False
            else:
                hit = asm.read_hit([str(x) for x in new[2:]], ["syn"], read.upper(), site_comp, st, et, 0, 0.2, 1, None)
                cols = range(hit.qstart, hit.qend + 1)
                case.update(keep=True, start_trim=st, end_trim=et, trimmed=[str(x) for x in new[2:6]], qstart=hit.qstart,
                            qend=hit.qend, base="".join(hit.base[c] for c in cols),
                            codon="".join("".join(list(hit.seq_comp[c][i].keys())[0] for i in (1, 2, 3)) for c in cols),
                            seqids=hit.seqids["syn"], unmap=sorted(hit.unmap["syn"].keys()))
            cases.append(case)
    json.dump(cases, open(os.path.join(HERE, "mapping_golden.json"), "w"), indent=0)
    print("mapping golden:", sum(1 for c in cases if "row" in c), "hits")


def main():
    os.makedirs(OUT, exist_ok=True)
    profiles = build_profiles()
    tmp = tempfile.mkdtemp(prefix="pa_golden_")
    try:
        names, seqs = run_reference_prepare_target(tmp)
    finally:
        pass
    print("targets:", len(names))
    allrows = oracle_hit_info(profiles, names, seqs)
    mapping_golden(allrows)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
