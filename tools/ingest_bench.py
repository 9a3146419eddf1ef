#!/usr/bin/env python
"""Native ingest vs the reference's prepare_target() on synthetic FASTQ (CPU only; the reference arm needs
/root/reference and the test-only Bio shim, so it runs in the build container)."""
import gzip, os, sys, tempfile, time
from argparse import Namespace
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phyloaln_b200 import ingest

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
rng = np.random.default_rng(7)
tmp = tempfile.mkdtemp()
files = []
for f in range(4):
    arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n // 4, 150))]
    dup = rng.random(n // 4) < 0.05
    arr[dup] = arr[0]
    p = os.path.join(tmp, f"r{f}.fq.gz")
    with gzip.open(p, "wb", compresslevel=1) as g:
        q = b"I" * 150
        g.write(b"".join(b"@read%d_%d 1:N:0\n%s\n+\n%s\n" % (f, i, arr[i].tobytes(), q) for i in range(n // 4)))
    files.append(p)
rawdata = {"SPA": files[:2], "SPB": files[2:]}
t0 = time.perf_counter()
ing = ingest.Ingest(rawdata, threads=os.cpu_count())
t1 = time.perf_counter()
print(f"native ingest : {n} reads in {t1 - t0:.2f} s = {n / (t1 - t0) / 1e6:.2f} M reads/s ({ing.n} unique records, {os.cpu_count()} threads)")
if os.path.isdir("/root/reference"):
    sys.path.insert(0, os.path.join(ROOT, "tests", "bio_shim")); sys.path.insert(0, "/root/reference"); sys.path.insert(0, "/root/reference/lib")
    import lib.aln as al
    cwd = os.getcwd(); os.chdir(tmp); os.makedirs("ok", exist_ok=True)
    args = Namespace(file_format="guess", split_len=None, split_slide=None, merge_len=250, mol_type="dna", gencode=1, no_reverse=True, codon=False, cpu=1)
    t0 = time.perf_counter(); al.prepare_target(rawdata, args); t2 = time.perf_counter() - t0
    same = open("all.raw.fasta", "rb").read() == ing.raw_fasta()
    os.chdir(cwd)
    print(f"reference     : {t2:.2f} s = {n / t2 / 1e6:.3f} M reads/s (prepare_target incl. its chunk copy); identical output: {same}")
