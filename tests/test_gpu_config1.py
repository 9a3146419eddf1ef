"""GPU parity at full config-1 size (BASELINE.json configs[0]): the reference's own test reads
(prepare_target output, committed as a golden fixture) against the five tests/ref profiles, and the
hit -> reference-column mapping against vectors produced by the reference's Python."""
import gzip
import json
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
C1 = os.path.join(GOLD, "config1")
GROUPS = ["OG0003531", "OG0003709", "OG0003820", "OG0003977", "OG0004212"]


def test_config1_dropin_outputs_identical_to_golden(tmp_path):
    from phyloaln_b200 import dropin
    out = tmp_path / "run"
    (out / "ref_hmm").mkdir(parents=True)
    for g in GROUPS:
        shutil.copy(os.path.join(C1, "ref_hmm", g + ".hmm"), out / "ref_hmm" / (g + ".hmm"))
    with gzip.open(os.path.join(C1, "all.raw.fasta.gz"), "rb") as f, open(out / "all.raw.fasta", "wb") as g:
        shutil.copyfileobj(f, g)
    counts = dropin.hmmsearch_step(str(out), GROUPS, mol_type="codon", gencode=1, write_text=True)
    total = 0
    import sys
    shim = os.path.join(os.path.dirname(__file__), "bio_shim")
    sys.path.insert(0, shim)
    try:
        from Bio import SearchIO
        for g in GROUPS:   # the `-o` text report carries the same rows (what the unmodified extract_reads would parse)
            want = [l.split("\t") for l in open(os.path.join(C1, "hit_info", g + ".hit_info.tsv")).read().splitlines()]
            got = [f for q in SearchIO.parse(str(out / "map" / (g + ".hmmsearch.txt")), "hmmer3-text") for hit in q for hsp in hit for f in hsp]
            assert [(str(f.query_start + 1), str(f.query_end), str(f.hit_start + 1), str(f.hit_end), str(f.query.seq).upper().replace(".", "-"),
                     str(f.hit.seq).upper().replace(".", "-")) for f in got] == [tuple(w[2:6] + w[8:10]) for w in want], g
            assert all(len(f.aln_annotation["PP"]) == len(str(f.hit.seq)) for f in got)
    finally:
        sys.path.remove(shim)
    for g in GROUPS:
        got = open(out / "map" / (g + ".hit_info.tsv")).read()
        want = open(os.path.join(C1, "hit_info", g + ".hit_info.tsv")).read()
        assert got == want, g
        assert os.path.isfile(out / "ok" / f"hmmsearch_{g}.ok") and os.path.isfile(out / "ok" / f"extract_reads_{g}.ok")
        # targets.fas: one record per distinct read, first-seen order, raw nucleotides
        ids = []
        for line in want.splitlines():
            rid = line.split("\t")[1]
            if rid not in ids:
                ids.append(rid)
        fas = open(out / "map" / (g + ".targets.fas")).read().splitlines()
        assert fas[0::2] == [">" + i for i in ids]
        total += counts[g]
    assert total == 777
    # second call: resume markers present -> nothing to do (lib/aln.py:653-654, 484-488)
    assert dropin.hmmsearch_step(str(out), GROUPS, mol_type="codon") == {}


@pytest.mark.parametrize("wait", [True, False])
def test_config1_from_fastq_through_native_ingest(tmp_path, wait):
    """fq.gz -> native ingest -> GPU search: same hit_info files as the golden run that started from the
    reference's all.raw.fasta.  wait=False: the search starts while the ingest is still decoding (streaming); config 1
    repeats read ids, so read_fastx's dictionary semantics are applied to the hits after the search."""
    from argparse import Namespace
    from phyloaln_b200 import dropin
    out = tmp_path / "run"
    (out / "ref_hmm").mkdir(parents=True)
    for g in GROUPS:
        shutil.copy(os.path.join(C1, "ref_hmm", g + ".hmm"), out / "ref_hmm" / (g + ".hmm"))
    rd = os.path.join(C1, "reads")
    rawdata = {"DMELA": [os.path.join(rd, "DMELA_1.fq.gz"), os.path.join(rd, "DMELA_2.fq.gz")],
               "DWILL": [os.path.join(rd, "DWILL_1.fq.gz"), os.path.join(rd, "DWILL_2.fq.gz")]}
    args = Namespace(file_format="guess", split_len=None, split_slide=None, merge_len=250, cpu=4)
    ing = dropin.prepare_target_step(str(out), rawdata, args, wait=wait)
    dropin.hmmsearch_step(str(out), GROUPS, mol_type="codon", gencode=1, ingest=ing, chunk_reads=2000)
    assert open(out / "total_counts.tsv").read() == open(os.path.join(C1, "total_counts.tsv")).read()
    assert gzip.open(os.path.join(C1, "all.raw.fasta.gz"), "rb").read() == open(out / "all.raw.fasta", "rb").read()
    assert os.path.isfile(out / "ok" / "prepare_target.ok") and os.path.isfile(out / "ok" / "prepare_target.b200")
    assert open(out / "all.target.fasta").read().startswith("#")    # a placeholder no sequence parser accepts
    for g in GROUPS:
        assert open(out / "map" / (g + ".hit_info.tsv")).read() == open(os.path.join(C1, "hit_info", g + ".hit_info.tsv")).read(), g


def test_config1_scores_within_tolerance(gpu_ctx_factory):
    from phyloaln_b200 import dropin, hmmio, search
    ids, nt, off = dropin.read_raw_fasta(os.path.join(C1, "all.raw.fasta.gz"))
    hmms = [hmmio.read_hmm(os.path.join(C1, "ref_hmm", g + ".hmm")) for g in GROUPS]
    s = search.B200Searcher(hmms)
    rep, nframes, translated = s.search_and_report(nt, off, ids, mol_type="codon")
    s.close()
    for hi, g in enumerate(GROUPS):
        want = [l.split("\t") for l in open(os.path.join(C1, "hit_info", g + ".scores.tsv")).read().splitlines()]
        assert len(want) == len(rep[hi])
        for w, r in zip(want, rep[hi]):
            assert w[0] == r["name"] and int(w[1]) == r["dom_idx"]
            assert abs(float(w[2]) - r["seq_bits"]) < 0.01 and abs(float(w[4]) - r["dom_bits"]) < 0.01
            # E-values within 1e-3 relative <=> lnP within 1e-3 absolute
            assert abs(float(w[3]) - np.log(r["seq_evalue"] / (len(ids) * nframes))) < 1e-3


def test_mapping_kernel_matches_reference_python(gpu_ctx_factory):
    from phyloaln_b200 import mapping
    cases = json.load(open(os.path.join(GOLD, "mapping_golden.json")))
    comps = {c["group"]: mapping.composition_table({int(k): v for k, v in c["site_composition"].items()})
             for c in cases if "site_composition" in c}
    kat = [c for c in cases if c["group"] == "KAT5"][0]
    comps["g"] = mapping.composition_table({int(k): v for k, v in kat["site_composition_inline"].items()})
    todo = [c for c in cases if "row" in c]
    rows, reads = [], {}
    for i, c in enumerate(todo):
        r = list(c["row"])
        r[1] = f"{r[1]}#{i}"           # unique key per case (several cases share a read id)
        reads[r[1]] = c["read"]
        rows.append(r)
    ctx = gpu_ctx_factory()
    got = mapping.map_hits(ctx, rows, reads, comps, trim_pos=5, pos_freq=0.2)
    ctx.close()
    n_trim = n_unmap = 0
    for c, g in zip(todo, got):
        if not c["keep"]:
            assert g is None
            continue
        assert g is not None, c["row"]
        assert (g["qstart"], g["qend"], g["start_trim"], g["end_trim"]) == (c["qstart"], c["qend"], c["start_trim"], c["end_trim"]), c["row"]
        assert [g["qstart"], g["qend"], g["ali_from"], g["ali_to"]] == [int(x) for x in c["trimmed"]]
        assert g["base"] == c["base"], c["row"]
        if c["codon"] is not None:
            assert g["codon"] == c["codon"], c["row"]
        assert [g["interval"]] == c["seqids"], c["row"]
        assert g["unmap"] == c["unmap"], c["row"]
        n_trim += bool(c["start_trim"] or c["end_trim"])
        n_unmap += bool(c["unmap"])
    assert n_trim > 20 and n_unmap > 20


def test_two_gpu_searcher_gives_the_one_gpu_report():
    """Reads shard by chunk over the GPUs of one box (one context per GPU, host-side gather, no collective): the report
    must not depend on the number of devices.  Skipped on a single-GPU box."""
    from phyloaln_b200 import capi, dropin, hmmio, search
    try:
        probe = capi.Context(1)
        probe.close()
    except capi.PaError:
        pytest.skip("needs two GPUs")
    ids, nt, off = dropin.read_raw_fasta(os.path.join(C1, "all.raw.fasta.gz"))
    hmms = [hmmio.read_hmm(os.path.join(C1, "ref_hmm", g + ".hmm")) for g in GROUPS]
    reports = []
    for devices in ((0,), (0, 1)):
        s = search.B200Searcher(hmms, devices=devices)
        rep, nframes, translated = s.search_and_report(nt, off, ids, mol_type="codon", chunk_reads=1500)
        s.close()
        reports.append([[(r["name"], r["dom_idx"], r["hmmfrom"], r["hmmto"], r["alifrom"], r["alito"], r["model"], r["aligned"],
                          round(r["seq_bits"], 3)) for r in rep[hi]] for hi in range(len(GROUPS))])
    assert reports[0] == reports[1] and sum(map(len, reports[0])) == 777


def test_config1_uniquified_ids_variant(oracle):
    """config 1 repeats read ids, which the reference's read_fastx dictionary collapses (first position, last
    sequence: 2,481 of 9,930 records survive -- the variant every other test pins).  The uniquified variant makes the harness
    append #n to repeated ids BEFORE both pipelines see them, so every record is a target of its own (nothing is dropped
    silently): GPU and oracle must then agree on all of them.  Records whose 'sequence' is not nucleotides (a FASTQ quirk of
    the reference's 4-line parser; Biopython would raise on them) are set aside explicitly and counted."""
    from phyloaln_b200 import hmmio, ingest, search
    rd = os.path.join(C1, "reads")
    rawdata = {"DMELA": [os.path.join(rd, "DMELA_1.fq.gz"), os.path.join(rd, "DMELA_2.fq.gz")],
               "DWILL": [os.path.join(rd, "DWILL_1.fq.gz"), os.path.join(rd, "DWILL_2.fq.gz")]}
    ing = ingest.Ingest(rawdata, merge_len=250, threads=4)
    ids, seen = [], {}
    for i in ing.ids():
        n = seen.get(i, 0)
        seen[i] = n + 1
        ids.append(i if n == 0 else f"{i}#{n}")
    assert len(set(ids)) == len(ids) == 9930
    nt, off = ing.pack()
    ing.close()
    valid = search.valid_reads_mask(nt, off)
    assert 0 < int((~valid).sum()) < 20                      # the junk records: reported, not dropped silently
    keep = np.flatnonzero(valid)
    raw = nt.tobytes().decode()
    reads = [raw[int(off[k]):int(off[k + 1])] for k in keep]
    kept_ids = [ids[k] for k in keep]
    sub_off = np.zeros(len(reads) + 1, dtype=np.uint64)
    sub_off[1:] = np.cumsum([len(r) for r in reads])
    sub_nt = np.frombuffer("".join(reads).encode(), dtype=np.uint8)
    hmms = [hmmio.read_hmm(os.path.join(C1, "ref_hmm", g + ".hmm")) for g in GROUPS]
    s = search.B200Searcher(hmms)
    rep, nframes, translated = s.search_and_report(sub_nt, sub_off, kept_ids, mol_type="codon", chunk_reads=4000)
    s.close()
    assert nframes == 6 and s.last_Z == 6 * len(reads)       # Z of the uniquified variant: every record counts
    frames = [p for r in reads for _, p in oracle.six_frames(r, 1)]
    names = [i + suf for i in kept_ids for suf in ("_pos1", "_pos2", "_pos3", "_rev_pos1", "_rev_pos2", "_rev_pos3")]
    from oracle import report as orep
    total = 0
    for hi, h in enumerate(hmms):
        ohits, _ = oracle.Profile.from_hmm(h).search(frames, nthreads=8, want_trace=False)
        want = orep.apply_report(ohits, names, len(frames))
        got = rep[hi]
        assert [(names[w["target"]], w["dom_idx"], w["hmmfrom"], w["hmmto"], w["iali"], w["jali"], w["model"], w["aligned"]) for w in want] == \
               [(g["name"], g["dom_idx"], g["hmmfrom"], g["hmmto"], g["alifrom"], g["alito"], g["model"], g["aligned"]) for g in got], GROUPS[hi]
        total += len(got)
    assert total > 777                                       # more rows than the dictionary variant: no record was collapsed
