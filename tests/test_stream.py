"""Host pipeline of the drop-in step without a GPU: streaming ingest (records visible while the files are still being
decoded), the chunk scheduler of B200Searcher.search_source (shared cursor, byte-bounded ranges, capacity splits), the
dictionary semantics of read_fastx applied after the search, and the vectorised report.  The GPU contexts are replaced by
a stand-in that "finds" a hit wherever a read starts with a marker -- the scheduling logic is what is under test."""
import gzip
import os
import threading
import time
from argparse import Namespace

import numpy as np
import pytest

from phyloaln_b200 import capi, dropin, ingest, search

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _config1_rawdata():
    rd = os.path.join(GOLD, "config1", "reads")
    return {"DMELA": [os.path.join(rd, "DMELA_1.fq.gz"), os.path.join(rd, "DMELA_2.fq.gz")],
            "DWILL": [os.path.join(rd, "DWILL_1.fq.gz"), os.path.join(rd, "DWILL_2.fq.gz")]}


def test_streaming_ingest_equals_blocking_ingest():
    a = ingest.Ingest(_config1_rawdata(), merge_len=250, threads=4)
    b = ingest.Ingest(_config1_rawdata(), merge_len=250, threads=4, wait=False)
    seen = 0
    while not b.done():
        n = b.emitted()
        assert n >= seen
        if n > seen:   # records already visible are final: same bytes as the finished ingest's
            nt, off = b.pack_emitted(seen, min(n - seen, 64))
            nt2, off2 = a.pack(seen, min(n - seen, 64))
            assert np.array_equal(off, off2) and nt[:int(off[-1])].tobytes() == nt2[:int(off2[-1])].tobytes()
        seen = n
        time.sleep(0.0005)
    b.wait()
    assert b.n == a.n == 9930 and b.emitted() == a.n
    assert b.raw_fasta() == a.raw_fasta() == gzip.open(os.path.join(GOLD, "config1", "all.raw.fasta.gz"), "rb").read()
    assert b.final_order() is None
    ids = a.ids()
    assert b.ids_of([0, 17, a.n - 1]) == [ids[0], ids[17], ids[-1]]
    recs = a.raw_fasta().decode().splitlines()
    assert b.seqs_of([3, 4000]) == [recs[2 * 3 + 1], recs[2 * 4000 + 1]]
    assert b.emitted_bytes(10, 5) == sum(len(recs[2 * k + 1]) for k in range(10, 15))
    # config 1 repeats read ids massively: 9,930 records, 2,481 distinct ids
    assert b.repeated_ids() == a.n - len(set(ids)) == 9930 - 2481
    a.close()
    b.close()


def test_ingest_error_surfaces_on_wait(tmp_path):
    bad = tmp_path / "x.fa"
    bad.write_text(">\nACGT\n")      # header without an identifier: the reference raises IndexError
    ing = ingest.Ingest({"S": [str(bad)]}, merge_len=250, file_format="fasta", wait=False)
    with pytest.raises(capi.PaError, match="identifier"):
        ing.wait()
    assert ing.done()
    ing.close()


class FakeCtx:
    """Stands in for capi.Context: one 'domain' for every read that starts with b'GATTACA'; -3 when a range is too big."""

    def __init__(self, max_reads=10 ** 9, delay=0.0):
        self.max_reads, self.delay, self.calls = max_reads, delay, []

    def load_reads(self, nt, off, moltype=0, gencode=1, no_reverse=False, codon_only=False):
        n = len(off) - 1
        if 2 * int(off[-1]) + 32 * n + 64 >= search.CHUNK_BYTE_LIMIT or n > self.max_reads:
            raise capi.PaError("pa_load_reads failed (-3): split the chunk")
        if not search.valid_reads_mask(nt, off).all():
            raise capi.PaError("pa_load_reads failed (-2): invalid nucleotide character in reads (Biopython would raise TranslationError)")
        self.nt, self.off = nt, off
        self.calls.append(n)
        return n * 6

    def num_frames(self):
        return 6

    def search(self):
        time.sleep(self.delay)
        raw = self.nt.tobytes()
        hit = [r for r in range(len(self.off) - 1) if raw[int(self.off[r]):int(self.off[r]) + 7] == b"GATTACA"]
        h = np.zeros(len(hit), dtype=capi.HIT_DTYPE)
        h["target"] = [6 * r + 2 for r in hit]
        h["seq_lnP"] = -30.0
        h["dom_lnP"] = -30.0
        return h, ["m"] * len(hit), ["a"] * len(hit)

    def hit_pp(self, hits):
        return ["9"] * len(hits)


def _searcher(ctxs):
    s = search.B200Searcher.__new__(search.B200Searcher)
    s.opts, s.hmms, s.ctxs, s._setup_lock, s._committed = search.SearchOptions(), [None], ctxs, threading.Lock(), True
    return s


def _reads(n, rng, marked):
    reads = [bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(30, 90))).astype(np.uint8)) for _ in range(n)]
    for r in marked:
        reads[r] = b"GATTACA" + reads[r]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    return np.frombuffer(b"".join(reads), dtype=np.uint8), off


def test_chunk_scheduler_global_indices_and_capacity_split():
    rng = np.random.default_rng(3)
    marked = [0, 5, 999, 1000, 2503, 4999]
    nt, off = _reads(5000, rng, marked)
    ctxs = [FakeCtx(max_reads=300), FakeCtx(max_reads=10 ** 9, delay=0.001)]
    hits, model, aligned, Z, nframes, translated = _searcher(ctxs).search_reads(nt, off, chunk_reads=1000, want_pp=True)
    assert (Z, nframes, translated) == (30000, 6, True)
    assert sorted(int(t) for t in hits["target"]) == [6 * r + 2 for r in marked]     # global target indices, every range once
    assert list(hits["target"]) == sorted(hits["target"])                           # ranges re-assembled in record order
    assert max(ctxs[0].calls, default=0) <= 300                                    # the -3 answers were bisected
    assert sum(ctxs[0].calls) + sum(ctxs[1].calls) == 5000


def test_scheduler_follows_a_source_that_is_still_producing():
    rng = np.random.default_rng(4)
    marked = [10, 700, 1500, 2999]
    nt, off = _reads(3000, rng, marked)

    class Trickle(search.ArraySource):
        def __init__(self):
            super().__init__(nt, off)
            self.n, self.fin = 0, False

        def available(self):
            return self.n

        def finished(self):
            return self.fin
    src = Trickle()

    def produce():
        for _ in range(30):
            time.sleep(0.002)
            src.n += 100
        src.fin = True
    t = threading.Thread(target=produce)
    t.start()
    ctxs = [FakeCtx(), FakeCtx()]
    hits, *_ , Z, nframes, _tr = _searcher(ctxs).search_source(src, chunk_reads=400)
    t.join()
    assert Z == 18000 and sorted(int(x) for x in hits["target"]) == [6 * r + 2 for r in marked]
    assert len(ctxs[0].calls) + len(ctxs[1].calls) >= 8 and max(ctxs[0].calls + ctxs[1].calls) <= 400


def test_single_read_beyond_the_chunk_limit_is_a_loud_error(monkeypatch):
    monkeypatch.setattr(search, "CHUNK_BYTE_LIMIT", 1000)
    nt, off = _reads(3, np.random.default_rng(1), [])
    big = np.frombuffer(b"A" * 600, dtype=np.uint8)
    nt = np.concatenate([nt, big])
    off = np.r_[off, off[-1] + 600].astype(np.uint64)
    with pytest.raises(capi.PaError):
        _searcher([FakeCtx()]).search_reads(nt, off, chunk_reads=2)


def test_dictionary_view_equals_read_fastx_semantics():
    """Hits of every emitted record -> the reference's dictionary view (first position, last sequence per id)."""
    ing = ingest.Ingest(_config1_rawdata(), merge_len=250, threads=2)
    ids_dict, nt_dict, off_dict = dropin.reads_from_ingest(ing)        # the plain statement of the semantics
    nframes = 6
    hits = np.zeros(ing.n, dtype=capi.HIT_DTYPE)                        # one fake domain per record, frame = record % 6
    hits["target"] = np.arange(ing.n) * nframes + np.arange(ing.n) % nframes
    hits["hmmfrom"] = np.arange(ing.n)
    out, sel, nent, entry_rec = dropin._dictionary_view(ing, hits, nframes)
    assert nent == len(ids_dict) == 2481
    all_ids = ing.ids()
    raw = nt_dict.tobytes()
    seqs = ing.seqs_of(np.arange(ing.n))
    for k in range(len(out)):
        rec = int(out["hmmfrom"][k])                                    # the record this surviving hit came from
        ent = int(out["target"][k]) // nframes
        assert ids_dict[ent] == all_ids[rec]
        assert raw[int(off_dict[ent]):int(off_dict[ent + 1])].decode() == seqs[rec]   # it is the id's LAST sequence
        assert int(out["target"][k]) % nframes == rec % nframes
    assert len(out) == nent and sorted(set(int(t) // nframes for t in out["target"])) == list(range(nent))
    ing.close()


def test_report_index_is_the_report(oracle):
    """report_index + rows_from_index == report (same rows, same order), with ties in E-value ordered by name."""
    rng = np.random.default_rng(8)
    n = 400
    hits = np.zeros(n, dtype=capi.HIT_DTYPE)
    hits["hmm"] = rng.integers(0, 3, n)
    hits["target"] = rng.integers(0, 120, n)
    key = hits["hmm"].astype(np.int64) * 1000 + hits["target"]
    for k in np.unique(key):                                 # consistent per-sequence fields, dom_idx 0..ndom-1
        m = np.flatnonzero(key == k)
        hits["dom_idx"][m] = np.arange(len(m))
        hits["ndom_in_seq"][m] = len(m)
        hits["seq_lnP"][m] = float(rng.choice([-2.0, -8.0, -8.0, -20.0, 1.5]))      # many ties
        hits["seq_bits"][m] = float(rng.normal(20, 5))
    hits["dom_lnP"] = rng.choice([-1.0, -6.0, 2.0], n)
    hits["dom_bits"] = rng.normal(10, 5, n)
    model = [f"m{i}" for i in range(n)]
    aligned = [f"a{i}" for i in range(n)]
    names = [f"read{(7 * t) % 120:03d}_pos1" for t in range(120)]
    for opts in (search.SearchOptions(), search.SearchOptions(E=1e-3, domE=1e-2), search.SearchOptions(T=18.0, domT=9.0),
                 search.SearchOptions(Z=5.0, domZ=2.0)):
        rep = search.report(hits, model, aligned, lambda t: names[t], 120, opts, 3)
        # reference statement: the per-hit loop this function replaced
        import math
        for h in range(3):
            groups = {}
            for i in np.lexsort((hits["dom_idx"], hits["target"], hits["hmm"])):
                if int(hits["hmm"][i]) == h:
                    groups.setdefault(int(hits["target"][i]), []).append(int(i))
            Zeff = opts.Z if opts.Z is not None else 120.0
            ok = [t for t, g in groups.items() if (hits["seq_bits"][g[0]] >= opts.T if opts.T is not None else math.exp(hits["seq_lnP"][g[0]]) * Zeff <= opts.E)]
            domZ = opts.domZ if opts.domZ is not None else float(len(ok))
            ok.sort(key=lambda t: (float(hits["seq_lnP"][groups[t][0]]), names[t].encode()))
            want = [i for t in ok for i in groups[t]
                    if (hits["dom_bits"][i] >= opts.domT if opts.domT is not None else math.exp(hits["dom_lnP"][i]) * domZ <= opts.domE)]
            assert [r["model"] for r in rep[h]] == [model[i] for i in want], (opts, h)
            assert all(r["domZ"] == domZ for r in rep[h])


def test_gather_hits_files_roundtrip(tmp_path):
    parts = []
    for rank in range(3):
        h = np.zeros(rank + 1, dtype=capi.HIT_DTYPE)
        h["target"] = 100 * rank + np.arange(rank + 1)
        parts.append((h, [f"m{rank}"] * (rank + 1), [f"a{rank}"] * (rank + 1)))
    res = {}

    def run(rank):
        res[rank] = search.gather_hits_files(*parts[rank], rank, 3, str(tmp_path / "g"), dst=0, timeout=30)
    th = [threading.Thread(target=run, args=(r,)) for r in (2, 1, 0)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert res[1] == (None, None, None) and res[2] == (None, None, None)
    hits, model, aligned = res[0]
    assert list(hits["target"]) == [0, 100, 101, 200, 201, 202] and model == ["m0", "m1", "m1", "m2", "m2", "m2"]
    assert os.listdir(tmp_path / "g") == []


def test_records_with_invalid_letters_are_set_aside_not_fatal():
    """A FASTQ quirk of the reference (config 1 has it) leaves records whose 'sequence' is a quality or header line; the
    reference's dictionary drops most of them before translating.  The scheduler searches the rest of the chunk, keeps
    global indices right and reports the records it left out."""
    rng = np.random.default_rng(9)
    marked = [3, 40, 41, 77]
    nt, off = _reads(100, rng, marked)
    nt = nt.copy()
    for r in (5, 41, 99):                        # 41 is a marked read: its hit must disappear with it
        nt[int(off[r]) + 2] = ord("!")
    s = _searcher([FakeCtx()])
    hits, *_ = s.search_reads(nt, off, chunk_reads=64)
    assert s.last_invalid == [5, 41, 99]
    assert sorted(int(t) for t in hits["target"]) == [6 * r + 2 for r in (3, 40, 77)]
    with pytest.raises(capi.PaError, match="invalid nucleotide"):
        s.search_and_report(nt, off, [f"r{i}" for i in range(100)], chunk_reads=64)
