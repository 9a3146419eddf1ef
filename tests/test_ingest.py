"""Native ingest (csrc/ingest.cpp) against the reference's own prepare_target() outputs
(tests/golden/make_ingest_golden.py, tests/golden/make_golden.py): byte-identical all.raw.fasta,
total_counts.tsv and merged_redundance.txt.  Host code only -- runs without a GPU."""
import glob
import gzip
import json
import os
from argparse import Namespace

import numpy as np
import pytest

from phyloaln_b200 import ingest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(GOLD, "ingest", "*", "case.json")))


def test_cases_exist():
    assert len(CASES) >= 7


@pytest.mark.parametrize("case", CASES)
def test_ingest_matches_reference_prepare_target(case, tmp_path):
    cdir = os.path.join(GOLD, "ingest", case)
    c = json.load(open(os.path.join(cdir, "case.json")))
    indir = os.path.join(GOLD, "ingest", c["indir"])
    rawdata = {sp: [os.path.join(indir, f) for f in fs] for sp, fs in c["rawdata"].items()}
    args = Namespace(cpu=3, **c["args"])
    ing = ingest.prepare_target(rawdata, args, outdir=str(tmp_path))
    for f in ("all.raw.fasta", "total_counts.tsv", "merged_redundance.txt"):
        assert open(os.path.join(tmp_path, f), "rb").read() == open(os.path.join(cdir, f), "rb").read(), (case, f)
    # record table is consistent with the text
    want = open(os.path.join(cdir, "all.raw.fasta")).read()
    ids = [line[1:] for line in want.splitlines() if line.startswith(">")]
    assert ing.ids() == ids
    ing.close()


def test_config1_reads_match_golden(tmp_path):
    """The reference's own test reads (tests/*.fq.gz) -> the committed prepare_target() output of config 1."""
    rd = os.path.join(GOLD, "config1", "reads")
    rawdata = {"DMELA": [os.path.join(rd, "DMELA_1.fq.gz"), os.path.join(rd, "DMELA_2.fq.gz")],
               "DWILL": [os.path.join(rd, "DWILL_1.fq.gz"), os.path.join(rd, "DWILL_2.fq.gz")]}
    ing = ingest.Ingest(rawdata, merge_len=250, threads=4)
    assert ing.raw_fasta() == gzip.open(os.path.join(GOLD, "config1", "all.raw.fasta.gz"), "rb").read()
    assert ing.total_counts() == open(os.path.join(GOLD, "config1", "total_counts.tsv")).read()
    assert ing.merged_redundance() == open(os.path.join(GOLD, "config1", "merged_redundance.txt")).read()
    assert ing.n == 9930
    # pack(): contiguous nucleotides + offsets, ready for pa_load_reads
    nt, off = ing.pack(100, 50)
    recs = ing.raw_fasta().decode().splitlines()
    for k in range(50):
        assert nt[int(off[k]):int(off[k + 1])].tobytes().decode() == recs[2 * (100 + k) + 1]
    ing.close()


def test_chunk_rotation_keeps_order(tmp_path):
    """More than 10 MB of sequence rotates chunk files in the reference; without the >1000-chunk fold the
    concatenation order is the write order.  Property: ids come out in input order, nothing is lost."""
    rng = np.random.default_rng(1)
    p = tmp_path / "big.fa"
    n = 1300
    with open(p, "w") as f:
        for i in range(n):
            f.write(f">s{i} x\n")
            s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 10000)).decode()
            for k in range(0, len(s), 80):
                f.write(s[k:k + 80] + "\n")
    ing = ingest.Ingest({"SP": [str(p)]}, merge_len=250, threads=2)
    assert ing.ids() == [f"s{i}_PhAlTag_SP_fastx1" for i in range(n)]
    assert (ing.seq_lengths() == 10000).all()
    assert ing.total_counts() == f"SP\t{n}\n"
    ing.close()


def test_bad_inputs_are_errors(tmp_path):
    from phyloaln_b200.capi import PaError
    with pytest.raises(PaError):
        ingest.Ingest({"SP": [str(tmp_path / "missing.fq.gz")]})
    with pytest.raises(PaError):
        ingest.Ingest({"SP": [__file__]}, file_format="sam")
