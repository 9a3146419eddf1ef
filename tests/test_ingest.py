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


def _golden_bytes(path):
    import gzip
    if os.path.exists(path):
        return open(path, "rb").read()
    return gzip.open(path + ".gz", "rb").read()


def test_cases_exist():
    assert len(CASES) >= 11 and sum(c.startswith("fold_") for c in CASES) == 4


@pytest.mark.parametrize("case", CASES)
def test_ingest_matches_reference_prepare_target(case, tmp_path, monkeypatch):
    cdir = os.path.join(GOLD, "ingest", case)
    c = json.load(open(os.path.join(cdir, "case.json")))
    indir = os.path.join(GOLD, "ingest", c["indir"])
    rawdata = {sp: [os.path.join(indir, f) for f in fs] for sp, fs in c["rawdata"].items()}
    args = Namespace(cpu=3, **c["args"])
    if "each_size" in c:     # the >1000-chunk fold cases: the reference ran with its 10 MB chunk constant replaced by this
        monkeypatch.setenv("PA_INGEST_EACH_SIZE", str(c["each_size"]))
    ing = ingest.prepare_target(rawdata, args, outdir=str(tmp_path))
    for f in ("all.raw.fasta", "total_counts.tsv", "merged_redundance.txt"):
        assert open(os.path.join(tmp_path, f), "rb").read() == _golden_bytes(os.path.join(cdir, f)), (case, f)
    # record table is consistent with the text
    want = _golden_bytes(os.path.join(cdir, "all.raw.fasta")).decode()
    ids = [line[1:] for line in want.splitlines() if line.startswith(">")]
    assert ing.ids() == ids
    if "each_size" in c:
        # the fold really happened: the file order is a non-trivial permutation of the emission order, and the ids the
        # search path reads in emission order (pa_ingest_emitted_ids) map onto the file through pa_ingest_final_order
        perm = ing.final_order()
        assert sorted(perm.tolist()) == list(range(len(ids))) and not (perm == np.arange(len(ids))).all()
        emitted = ing.ids_of(np.arange(len(ids)))
        assert [emitted[int(k)] for k in perm] == ids
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


def test_native_raw_fasta_loader_equals_python_statement(tmp_path):
    """pa_fasta_* == dropin.read_raw_fasta (the plain-Python statement of read_fastx's dictionary semantics,
    lib/library.py:186-198): repeated ids (first position, last sequence), multi-line records, CRLF, blank lines, text
    before the first header, header descriptions, gz input, an empty file, and the committed config-1 all.raw.fasta."""
    import gzip
    from phyloaln_b200 import dropin
    from phyloaln_b200.ingest import load_raw_fasta
    text = ("junk before\n>r1 desc here\nACGT\n>r2\nAC\nGT\r\n\nTT\n>r1\nGGGG\n>r3\n>r4_PhAlTag_SP_fastx1\nacgtn\n>r2\nC\n>\nAAA\n"
            ">last")
    p = tmp_path / "a.fasta"
    p.write_bytes(text.encode())
    with gzip.open(tmp_path / "a.fasta.gz", "wb") as f:
        f.write(text.encode())
    (tmp_path / "empty.fasta").write_bytes(b"")
    gold = os.path.join(os.path.dirname(__file__), "golden", "config1", "all.raw.fasta.gz")
    for path in (str(p), str(tmp_path / "a.fasta.gz"), str(tmp_path / "empty.fasta"), gold):
        ids0, nt0, off0 = dropin.read_raw_fasta(path)
        ids1, nt1, off1 = load_raw_fasta(path)
        assert list(ids1) == ids0 and len(ids1) == len(ids0)
        assert np.array_equal(off1, off0) and nt1.tobytes() == nt0.tobytes()
    ids, nt, off = load_raw_fasta(str(p))
    assert list(ids) == ["r1", "r2", "r3", "r4_PhAlTag_SP_fastx1", "", "last"]
    assert [nt.tobytes()[int(off[i]):int(off[i + 1])].decode() for i in range(len(ids))] == ["GGGG", "C", "", "acgtn", "AAA", ""]
    assert ids[-1] == "last" and ids[1:3] == ["r2", "r3"]
    with pytest.raises(OSError):
        load_raw_fasta(str(tmp_path / "missing.fasta"))
    # many tiny records (the id table grows several times) with every 7th id repeated later
    names = [f"s{i}" for i in range(6000)] + [f"s{i}" for i in range(0, 6000, 7)]
    (tmp_path / "tiny.fasta").write_text("".join(f">{nm}\n{'ACGT'[k % 4]}\n" for k, nm in enumerate(names)))
    ids0, nt0, off0 = dropin.read_raw_fasta(str(tmp_path / "tiny.fasta"))
    ids1, nt1, off1 = load_raw_fasta(str(tmp_path / "tiny.fasta"))
    assert list(ids1) == ids0 and len(ids0) == 6000 and nt1.tobytes() == nt0.tobytes() and np.array_equal(off1, off0)


@pytest.mark.parametrize("threads", [1, 2, 3, 7, 16])
def test_native_raw_fasta_loader_parallel_ranges(tmp_path, monkeypatch, threads):
    """The loader parses the file in byte ranges cut at header lines, one per thread, and applies read_fastx's dictionary in a
    serial pass over the ranges: the result must not depend on where the cuts fall -- repeated ids on both sides of a cut,
    multi-line and empty sequences, CRLF, '>' only at line starts, text before the first header, a file without trailing
    newline.  Random files against the plain-Python statement, for several thread counts."""
    from phyloaln_b200 import dropin
    from phyloaln_b200.ingest import load_raw_fasta
    monkeypatch.setenv("PA_INGEST_THREADS", str(threads))
    rng = np.random.default_rng(100 + threads)
    for case in range(12):
        nrec = int(rng.integers(1, 400))
        names = [f"id{int(rng.integers(0, max(2, nrec // 2)))}" for _ in range(nrec)]       # about half the ids repeat
        out = ["leading text\n"] if case % 3 == 0 else []
        for nm in names:
            out.append(">" + nm + (" description > with a bracket" if rng.random() < 0.3 else "") + ("\r\n" if rng.random() < 0.1 else "\n"))
            for _ in range(int(rng.integers(0, 4))):
                line = "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, int(rng.integers(0, 30))))
                out.append(line + ("\r\n" if rng.random() < 0.1 else "\n"))
        text = "".join(out)
        if case % 4 == 1:
            text = text.rstrip("\r\n")
        p = tmp_path / f"c{case}.fasta"
        p.write_bytes(text.encode())
        ids0, nt0, off0 = dropin.read_raw_fasta(str(p))
        ids1, nt1, off1 = load_raw_fasta(str(p))
        assert list(ids1) == ids0, (threads, case)
        assert np.array_equal(off1, off0) and nt1.tobytes() == nt0.tobytes(), (threads, case)
