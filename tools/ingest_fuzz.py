#!/usr/bin/env python
"""Differential fuzz of the native ingest against the reference's prepare_target() (build container only: needs
/root/reference and the test-only Bio shim).  Random FASTQ / FASTA inputs with the quirks the parsers disagree on most easily
(invalid or empty headers, '@' quality lines, missing trailing lines, CRLF, blanks in ids, duplicates, empty reads, several
files and species), random merge_len / split settings; all.raw.fasta, total_counts.tsv and merged_redundance.txt must be
byte-identical.  usage: tools/ingest_fuzz.py [cases] [seed]"""
import gzip, os, shutil, sys, tempfile
from argparse import Namespace
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "bio_shim")); sys.path.insert(0, REF); sys.path.insert(0, os.path.join(REF, "lib"))
import contextlib, io
import lib.aln as al
from phyloaln_b200 import ingest

def rnd_seq(rng, n):
    return "".join("ACGTN"[i] for i in rng.choice(5, n, p=[.24, .24, .24, .24, .04]))

def fastq(rng, n, pool):
    out = []
    for k in range(n):
        r = rng.random()
        name = f"r{k}" + (" 1:N:0 x" if rng.random() < 0.3 else "") + ("/1" if rng.random() < 0.2 else "")
        seq = pool[int(rng.integers(len(pool)))] if rng.random() < 0.3 else rnd_seq(rng, int(rng.integers(0, 70)))
        hdr = "@" + name
        if r < 0.03: hdr = "@"                      # empty id
        elif r < 0.06: hdr = name                   # no '@'
        elif r < 0.08: hdr = "@ " + name            # id starting with a blank
        elif r < 0.10: hdr = ""                     # empty header line
        qual = ("@" + "I" * (len(seq) - 1)) if seq and rng.random() < 0.2 else "I" * len(seq)
        out += [hdr, seq if rng.random() > 0.02 else seq.lower(), "+", qual]
    if rng.random() < 0.3: out = out[:-int(rng.integers(1, 4))]        # truncated last record
    nl = "\r\n" if rng.random() < 0.2 else "\n"
    return nl.join(out) + (nl if rng.random() < 0.8 else "")

def fasta(rng, n, pool):
    out = []
    for k in range(n):
        seq = pool[int(rng.integers(len(pool)))] if rng.random() < 0.3 else rnd_seq(rng, int(rng.integers(0, 200)))
        out.append(">" + (">" if rng.random() < 0.05 else "") + f"c{k}" + (" desc here" if rng.random() < 0.4 else ""))
        w = int(rng.integers(20, 80))
        out += [seq[i:i + w] for i in range(0, len(seq), w)]
    return "\n".join(out) + "\n"

def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    for c in range(ncases):
        tmp = tempfile.mkdtemp()
        pool = [rnd_seq(rng, int(n)) for n in rng.integers(10, 60, 12)]
        fmt = "fastq" if rng.random() < 0.7 else "fasta"
        rawdata = {}
        for s in range(int(rng.integers(1, 4))):
            files = []
            for j in range(int(rng.integers(1, 4))):
                text = (fastq if fmt == "fastq" else fasta)(rng, int(rng.integers(0, 60)), pool)
                gz = rng.random() < 0.5
                p = os.path.join(tmp, f"s{s}_{j}.{'fq' if fmt == 'fastq' else 'fa'}" + (".gz" if gz else ""))
                with (gzip.open(p, "wb") if gz else open(p, "wb")) as fh:     # closed before anyone reads it
                    fh.write(text.encode())
                files.append(p)
            rawdata[f"SP{s}"] = files
        split = None if rng.random() < 0.7 else int(rng.integers(8, 40))
        args = dict(file_format="guess" if rng.random() < 0.5 else fmt, split_len=split,
                    split_slide=None if split is None else int(rng.integers(1, split + 1)), merge_len=int(rng.choice([0, 20, 40, 250])))
        if fmt == "fasta" and any(os.path.getsize(f) == 0 for fs in rawdata.values() for f in fs) and args["file_format"] == "guess":
            args["file_format"] = fmt
        want = None
        cwd = os.getcwd(); ref = tempfile.mkdtemp(); os.chdir(ref); os.makedirs("ok")
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                al.prepare_target(rawdata, Namespace(mol_type="dna", gencode=1, no_reverse=True, codon=False, cpu=1, **args))
            want = [open(f, "rb").read() for f in ("all.raw.fasta", "total_counts.tsv", "merged_redundance.txt")]
        except Exception as e:
            want = type(e).__name__
        finally:
            os.chdir(cwd)
        out = tempfile.mkdtemp()
        try:
            ing = ingest.prepare_target(rawdata, Namespace(cpu=3, **args), outdir=out)
            ing.close()
            got = [open(os.path.join(out, f), "rb").read() for f in ("all.raw.fasta", "total_counts.tsv", "merged_redundance.txt")]
        except Exception as e:
            got = type(e).__name__
        same = (isinstance(want, str) and isinstance(got, str)) or want == got
        if not same:
            bad += 1
            if not isinstance(want, str) and not isinstance(got, str):
                for name, w, g in zip(("all.raw.fasta", "total_counts.tsv", "merged_redundance.txt"), want, got):
                    if w != g:
                        wl, gl = w.decode().split("\n"), g.decode().split("\n")
                        i = next((i for i, (a, b) in enumerate(zip(wl, gl)) if a != b), min(len(wl), len(gl)))
                        print(f"  {name}: {len(wl)} vs {len(gl)} lines; line {i}: ref {wl[i][:100] if i < len(wl) else None!r} / native {gl[i][:100] if i < len(gl) else None!r}")
            print("MISMATCH case", c, args, {k: [os.path.basename(x) for x in v] for k, v in rawdata.items()}, "reference:", want if isinstance(want, str) else "ok", "native:", got if isinstance(got, str) else "ok", "kept in", tmp)
        else:
            shutil.rmtree(tmp)
        shutil.rmtree(ref); shutil.rmtree(out)
    print(f"{ncases} cases, {bad} mismatches")
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
