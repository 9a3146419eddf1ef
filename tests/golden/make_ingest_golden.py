#!/usr/bin/env python
"""Golden vectors for the native ingest (phyloaln_b200/csrc/ingest.cpp), produced by the REFERENCE's own
prepare_target() (lib/aln.py:274-433) imported from /root/reference with the test-only Bio shim.
Run in the build container only.  Inputs are written next to the outputs so the tests travel:
  tests/golden/ingest/<case>/in/*            synthetic FASTQ/FASTA(.gz) inputs
  tests/golden/ingest/<case>/case.json       species -> files, arguments
  tests/golden/ingest/<case>/all.raw.fasta, total_counts.tsv, merged_redundance.txt   reference outputs
tests/golden/config1/reads/*.fq.gz are the reference's own test reads (tests/*.fq.gz, data only).
"""
import gzip
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


def rnd_seq(rng, n, alphabet="ACGT"):
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))


def write(path, text, gz=False, newline="\n"):
    data = text.replace("\n", newline).encode()
    if gz:
        with gzip.GzipFile(path, "wb", mtime=0) as g:
            g.write(data)
    else:
        open(path, "wb").write(data)


def make_cases(base):
    rng = np.random.default_rng(20261020)
    cases = {}

    # ---- fastq quirks: ids with spaces, '@' quality lines, duplicates within/between species, long reads, empty read
    d = os.path.join(base, "fastq_quirks", "in")
    os.makedirs(d)
    pool = [rnd_seq(rng, int(n)) for n in rng.integers(20, 60, 40)]
    def fq(recs):
        return "".join(f"@{i}\n{s}\n+\n{'@' + 'I' * (len(s) - 1) if s else ''}\n" for i, s in recs)
    a = [(f"readA{k} 1:N:0:{k} extra field", pool[int(rng.integers(len(pool)))]) for k in range(60)]
    a[7] = ("empty read", "")
    a[9] = ("lower", pool[3].lower())
    a[11] = ("withN", pool[4][:10] + "NNNN" + pool[4][14:])
    b = [(f"readB{k}/2", pool[int(rng.integers(len(pool)))]) for k in range(50)]
    b.append(("long1", rnd_seq(rng, 300)))
    b.append(("long1dup", b[-1][1]))
    c = [(f"c{k}", pool[int(rng.integers(len(pool)))]) for k in range(30)]
    write(os.path.join(d, "A_1.fq.gz"), fq(a), gz=True)
    write(os.path.join(d, "A_2.fq"), fq(b), newline="\r\n")
    write(os.path.join(d, "B_1.fastq.gz"), fq(c), gz=True)
    cases["fastq_quirks"] = dict(rawdata={"SPA": ["A_1.fq.gz", "A_2.fq"], "SPB": ["B_1.fastq.gz"]},
                                 args=dict(file_format="guess", split_len=None, split_slide=None, merge_len=250))
    cases["fastq_mergelen40"] = dict(rawdata={"SPA": ["A_1.fq.gz", "A_2.fq"], "SPB": ["B_1.fastq.gz"]},
                                     args=dict(file_format="fastq", split_len=None, split_slide=None, merge_len=40), share="fastq_quirks")

    # ---- multi-line FASTA, descriptions, duplicates (incl. the last record of a file), >> headers
    d = os.path.join(base, "fasta_multi", "in")
    os.makedirs(d)
    seqs = [rnd_seq(rng, int(n)) for n in rng.integers(30, 400, 25)]
    def fa(recs, width=60):
        out = []
        for i, s in recs:
            out.append(f">{i}\n")
            out += [s[k:k + width] + "\n" for k in range(0, len(s), width)]
        return "".join(out)
    r1 = [(f"tr{k} len={len(seqs[k])} some description", seqs[k]) for k in range(15)]
    r1.append(("dupOfTr2\tTabbed", seqs[2]))
    r1.append((">doubleGt", seqs[5][:50]))
    r1.append(("lastDup", seqs[3]))                     # last record, duplicate -> untagged id bug
    r2 = [(f"u{k}", seqs[10 + k]) for k in range(12)]
    r2.append(("emptyseq", ""))
    r2.append(("mixedCase", seqs[1][:40].lower() + seqs[1][40:80]))
    write(os.path.join(d, "asm1.fa"), fa(r1))
    write(os.path.join(d, "asm2.fasta.gz"), fa(r2, width=70), gz=True)
    cases["fasta_multi"] = dict(rawdata={"SP1": ["asm1.fa"], "SP2": ["asm2.fasta.gz", "asm1.fa"]},
                                args=dict(file_format="guess", split_len=None, split_slide=None, merge_len=250))
    cases["fasta_fastpath"] = dict(rawdata={"SP1": ["asm1.fa"], "SP2": ["asm2.fasta.gz", "asm1.fa"]},
                                   args=dict(file_format="guess", split_len=None, split_slide=None, merge_len=0), share="fasta_multi")
    cases["fasta_split"] = dict(rawdata={"SP1": ["asm1.fa"], "SP2": ["asm2.fasta.gz"]},
                                args=dict(file_format="fasta", split_len=50, split_slide=25, merge_len=250), share="fasta_multi")
    cases["fasta_split_odd"] = dict(rawdata={"SP1": ["asm1.fa"], "SP2": ["asm2.fasta.gz"]},
                                    args=dict(file_format="guess", split_len=33, split_slide=10, merge_len=20), share="fasta_multi")
    cases["fastq_split"] = dict(rawdata={"SPA": ["A_1.fq.gz", "A_2.fq"], "SPB": ["B_1.fastq.gz"]},
                                args=dict(file_format="guess", split_len=16, split_slide=8, merge_len=250), share="fastq_quirks")

    # ---- the >1000-chunk fold (check_file_num, lib/aln.py:261-271): the reference starts a new chunk file every 10 MB of
    # sequence and, at the 1000th, appends files 100..999 to files 0..99 and multiplies the chunk size by 10, so the final
    # concatenation is no longer in emission order.  Reaching that with real sizes takes 10 GB of reads; here the SAME code runs
    # with its constant `each_size = 10000000` replaced by a few bytes (run_reference(each_size=...)); the native ingest takes the
    # same number through PA_INGEST_EACH_SIZE.  Up to three folds per case.
    d = os.path.join(base, "fold_fastq", "in")
    os.makedirs(d)
    pool = [rnd_seq(rng, int(n)) for n in rng.integers(20, 41, 900)]
    def reads(prefix, n):
        return [(f"{prefix}{k}", pool[int(rng.integers(len(pool)))] if rng.random() < 0.25 else rnd_seq(rng, int(rng.integers(20, 41)))) for k in range(n)]
    write(os.path.join(d, "X_1.fq.gz"), fq(reads("x", 3200)), gz=True)
    write(os.path.join(d, "X_2.fq.gz"), fq(reads("y", 2400)), gz=True)
    write(os.path.join(d, "Y_1.fq"), fq(reads("z", 1900)))
    cases["fold_fastq"] = dict(rawdata={"SX": ["X_1.fq.gz", "X_2.fq.gz"], "SY": ["Y_1.fq"]}, each_size=1,
                               args=dict(file_format="guess", split_len=None, split_slide=None, merge_len=250))
    d = os.path.join(base, "fold_fasta", "in")
    os.makedirs(d)
    contigs = [(f"ctg{k} len", rnd_seq(rng, int(rng.integers(30, 90)))) for k in range(1500)]
    for k in range(0, 1500, 7):
        contigs[k] = (contigs[k][0], contigs[(k * 3) % 1500][1])          # duplicates: merge groups, no size added
    write(os.path.join(d, "asm.fa"), fa(contigs, width=40))
    cases["fold_fasta_split"] = dict(rawdata={"SP": ["asm.fa"]}, each_size=50, share="fold_fasta",
                                     args=dict(file_format="fasta", split_len=16, split_slide=8, merge_len=250))
    cases["fold_fasta_fastpath"] = dict(rawdata={"SP": ["asm.fa"], "SQ": ["asm.fa"]}, each_size=30, share="fold_fasta",
                                        args=dict(file_format="guess", split_len=None, split_slide=None, merge_len=0))
    cases["fold_fasta_merge"] = dict(rawdata={"SP": ["asm.fa"], "SQ": ["asm.fa"]}, each_size=64, share="fold_fasta",
                                     args=dict(file_format="guess", split_len=None, split_slide=None, merge_len=60))
    return cases


def reference_aln(each_size=None):
    """The reference's lib/aln.py as a module; with each_size, the same source with its chunk-size constant replaced."""
    if each_size is None:
        import lib.aln as al
        return al
    import types
    path = os.path.join(REF, "lib", "aln.py")
    src = open(path).read()
    assert src.count("each_size = 10000000") == 1
    mod = types.ModuleType(f"aln_each_size_{each_size}")
    mod.__file__ = path
    exec(compile(src.replace("each_size = 10000000", f"each_size = {int(each_size)}"), path, "exec"), mod.__dict__)
    return mod


def run_reference(indir, rawdata, args, outdir, each_size=None):
    al = reference_aln(each_size)
    ns = Namespace(mol_type="dna", gencode=1, no_reverse=True, codon=False, cpu=1, **args)
    rd = {sp: [os.path.join(indir, f) for f in fs] for sp, fs in rawdata.items()}
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    os.makedirs("ok", exist_ok=True)
    try:
        al.prepare_target(rd, ns)
    finally:
        os.chdir(cwd)
    for f in ("all.raw.fasta", "total_counts.tsv", "merged_redundance.txt"):
        shutil.copy(os.path.join(tmp, f), os.path.join(outdir, f))
    shutil.rmtree(tmp)


def main():
    base = os.path.join(HERE, "ingest")
    shutil.rmtree(base, ignore_errors=True)
    os.makedirs(base)
    cases = make_cases(base)
    for name, c in cases.items():
        cdir = os.path.join(base, name)
        os.makedirs(cdir, exist_ok=True)
        indir = os.path.join(base, c.get("share", name), "in")
        run_reference(indir, c["rawdata"], c["args"], cdir, c.get("each_size"))
        meta = dict(rawdata=c["rawdata"], args=c["args"], indir=os.path.join(c.get("share", name), "in"))
        if c.get("each_size"):
            meta["each_size"] = c["each_size"]
        json.dump(meta, open(os.path.join(cdir, "case.json"), "w"), indent=1)
        size = os.path.getsize(os.path.join(cdir, "all.raw.fasta"))
        if c.get("each_size"):                     # the fold cases are larger: keep their raw fasta compressed
            data = open(os.path.join(cdir, "all.raw.fasta"), "rb").read()
            with gzip.GzipFile(os.path.join(cdir, "all.raw.fasta.gz"), "wb", mtime=0) as g:
                g.write(data)
            os.remove(os.path.join(cdir, "all.raw.fasta"))
        print(name, size, "bytes")


if __name__ == "__main__":
    main()
