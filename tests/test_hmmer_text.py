"""HMMER3 text report emitter: the report written from hit records must give back, through a
Bio.SearchIO-style 'hmmer3-text' parse (tests/bio_shim) and the reference's own extract_reads()
(/root/reference/lib/aln.py:482-570, when the reference tree is present), exactly the hit_info rows the batched
drop-in writes directly."""
import io
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "bio_shim")
REF_LIB = "/root/reference/lib"


def _report_inputs(oracle):
    from phyloaln_b200 import capi, search
    from tests import helpers
    rng = np.random.default_rng(5)
    hmm = helpers.calibrated_hmm("OGtext", 150, seed=91, bg_mix=0.1)
    seqs = helpers.protein_targets([hmm], 400, rng, Lmin=40, Lmax=200, plant_every=3)
    names = [f"read{i // 6}_PhAlTag_SP_fastx1" + ("_rev" if i % 6 >= 3 else "") + f"_pos{i % 3 + 1}" for i in range(len(seqs))]
    oh, _ = oracle.Profile.from_hmm(hmm).search(seqs, nthreads=4, want_trace=False)
    recs, model, aligned, pp = [], [], [], []
    for d in oh:
        r = np.zeros(1, dtype=capi.HIT_DTYPE)[0]
        for f in ("target", "ndom_in_seq", "dom_idx", "ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto", "dom_bits", "dom_lnP",
                  "seq_bits", "seq_lnP", "seqbias", "dombias", "oasc"):
            r[f] = d[f]
        recs.append(r)
        model.append(d["model"])
        aligned.append(d["aligned"])
        pp.append(d["pp"])
    rows = search.report(np.array(recs, dtype=capi.HIT_DTYPE), model, aligned, lambda t: names[t], len(seqs), search.SearchOptions(),
                         1, pp=pp)[0]
    return hmm, seqs, names, rows


def _direct_rows(hmm, rows):
    out = []
    for r in rows:
        hit_id, strand, start_pos = r["name"], "+", None
        start_pos = int(hit_id[-1]) - 1
        hit_id = hit_id[:-5]
        if hit_id.endswith("_rev"):
            hit_id, strand = hit_id[:-4], "rev"
        out.append("\t".join(str(x) for x in (hmm.name, hit_id, r["hmmfrom"], r["hmmto"], r["alifrom"], r["alito"], strand, start_pos,
                                              r["model"].upper().replace(".", "-"), r["aligned"].upper().replace(".", "-"))))
    return out


@pytest.mark.parametrize("textw", [120, 90, None])
def test_text_report_round_trips_through_a_searchio_parse(oracle, textw):
    from phyloaln_b200 import hmmer_text
    hmm, seqs, names, rows = _report_inputs(oracle)
    assert len(rows) > 30 and any("." in r["model"] for r in rows) and any("-" in r["aligned"] for r in rows)
    buf = io.StringIO()
    hmmer_text.write_text_report(buf, hmm, rows, len(seqs), target_lengths=lambda t: len(seqs[t]), textw=textw)
    text = buf.getvalue()
    assert text.startswith("# hmmsearch ::") and text.rstrip().endswith("[ok]") and f"Query:       {hmm.name}  [M={hmm.M}]" in text
    if textw:
        assert max(len(l) for l in text.splitlines() if l.rstrip().endswith(" PP")) <= textw
    sys.path.insert(0, SHIM)
    try:
        from Bio import SearchIO
        got = []
        for q in SearchIO.parse(io.StringIO(text), "hmmer3-text"):
            for hit in q:
                for hsp in hit:
                    for f in hsp:
                        got.append((f.query_id, f.hit_id, f.query_start + 1, f.query_end, f.hit_start + 1, f.hit_end,
                                    str(f.query.seq), str(f.hit.seq), f.aln_annotation["PP"]))
    finally:
        sys.path.remove(SHIM)
    want = [(hmm.name, r["name"], r["hmmfrom"], r["hmmto"], r["alifrom"], r["alito"], r["model"], r["aligned"], r["pp"]) for r in rows]
    assert got == want
    # the match line: identities carry the model letter, the rest is '+' or blank, never a letter on a mismatch
    for r in rows[:20]:
        ml = hmmer_text.match_line(hmm, r["model"], r["aligned"], r["hmmfrom"])
        assert len(ml) == len(r["model"])
        for m, a, c in zip(r["model"], r["aligned"], ml):
            assert (c == m) if (m != "." and a != "-" and m.upper() == a.upper()) else c in "+ "


def test_domtblout_columns(oracle, tmp_path):
    from phyloaln_b200 import search
    hmm, seqs, names, rows = _report_inputs(oracle)
    search.write_domtblout(str(tmp_path / "d.tbl"), hmm.name, hmm.M, rows, target_lengths=lambda t: len(seqs[t]))
    lines = [l.split() for l in open(tmp_path / "d.tbl") if not l.startswith("#")]
    assert len(lines) == len(rows) and all(len(l) == 23 for l in lines)
    for l, r in zip(lines, rows):
        assert (l[0], l[3], int(l[5]), int(l[15]), int(l[16]), int(l[17]), int(l[18]), int(l[19]), int(l[20])) == (
            r["name"], hmm.name, hmm.M, r["hmmfrom"], r["hmmto"], r["alifrom"], r["alito"], r["envfrom"], r["envto"])
        assert int(l[2]) == len(seqs[r["target"]]) and 0.0 <= float(l[21]) <= 1.0


def test_empty_report(oracle):
    from phyloaln_b200 import hmmer_text
    hmm, _, _, _ = _report_inputs(oracle)
    buf = io.StringIO()
    hmmer_text.write_text_report(buf, hmm, [], 10)
    assert "[No hits detected that satisfy reporting thresholds]" in buf.getvalue()
    sys.path.insert(0, SHIM)
    try:
        from Bio import SearchIO
        res = list(SearchIO.parse(io.StringIO(buf.getvalue()), "hmmer3-text"))
    finally:
        sys.path.remove(SHIM)
    assert len(res) == 1 and len(res[0]) == 0


@pytest.mark.skipif(not os.path.isdir(REF_LIB), reason="reference tree not present (GPU box)")
def test_unmodified_extract_reads_consumes_the_report(oracle, tmp_path, monkeypatch):
    """The reference's extract_reads() run on our report writes the same hit_info.tsv / targets.fas as the drop-in."""
    from phyloaln_b200 import hmmer_text
    hmm, seqs, names, rows = _report_inputs(oracle)
    monkeypatch.chdir(tmp_path)
    os.mkdir("map"), os.mkdir("ok")
    with open("map/OGtext.hmmsearch.txt", "w") as fh:
        hmmer_text.write_text_report(fh, hmm, rows, len(seqs), target_lengths=lambda t: len(seqs[t]))
    raw_ids = sorted({n.split("_rev")[0].rsplit("_pos", 1)[0] for n in names})
    with open("all.raw.fasta", "w") as fh:
        for i in raw_ids:
            fh.write(f">{i}\nACGTACGT\n")
    monkeypatch.syspath_prepend(REF_LIB)
    monkeypatch.syspath_prepend(SHIM)
    for m in ("aln", "library", "assemble"):
        sys.modules.pop(m, None)
    import aln
    aln.extract_reads("map/OGtext.hmmsearch.txt", "extract_reads_OGtext", search_mode="hmmer")
    assert open("map/OGtext.hit_info.tsv").read().splitlines() == _direct_rows(hmm, rows)
    fas = open("map/OGtext.targets.fas").read().splitlines()
    seen = []
    for l in _direct_rows(hmm, rows):
        if l.split("\t")[1] not in seen:
            seen.append(l.split("\t")[1])
    assert fas[0::2] == [">" + s for s in seen] and os.path.isfile("ok/extract_reads_OGtext.ok")


@pytest.mark.gpu
def test_hmmsearch_front_end_on_the_gpu(tmp_path, oracle):
    """bin/hmmsearch with the reference's command line (lib/aln.py:664-667): report == oracle rows, PP lines included."""
    import subprocess
    from oracle import report as orep
    from phyloaln_b200 import hmmio
    hmm, seqs, names, _ = _report_inputs(oracle)
    hmmio.write_hmm(hmm, str(tmp_path / "g.hmm"))
    with open(tmp_path / "t.fasta", "w") as fh:
        for n, s in zip(names, seqs):
            fh.write(f">{n}\n{s}\n")
    exe = os.path.join(os.path.dirname(HERE), "phyloaln_b200", "bin", "hmmsearch")
    rc = subprocess.run([sys.executable, exe, "-o", str(tmp_path / "g.txt"), "--tblout", str(tmp_path / "g.tbl"), "--cpu", "2",
                         str(tmp_path / "g.hmm"), str(tmp_path / "t.fasta")], capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr
    oh, _ = oracle.Profile.from_hmm(hmmio.read_hmm(str(tmp_path / "g.hmm"))).search(seqs, nthreads=4, want_trace=False)
    want = orep.apply_report(oh, names, len(seqs))
    sys.path.insert(0, SHIM)
    try:
        from Bio import SearchIO
        got = [(f.hit_id, f.query_start + 1, f.query_end, f.hit_start + 1, f.hit_end, str(f.query.seq), str(f.hit.seq), f.aln_annotation["PP"])
               for q in SearchIO.parse(str(tmp_path / "g.txt"), "hmmer3-text") for hit in q for hsp in hit for f in hsp]
    finally:
        sys.path.remove(SHIM)
    assert got == [(names[w["target"]], w["hmmfrom"], w["hmmto"], w["iali"], w["jali"], w["model"], w["aligned"], w["pp"]) for w in want]
    assert len(got) > 30
    # an option that cannot be honoured -> exit status 1 (the reference then reports "Error in hmmsearch command"):
    # --cut_ga on a profile without GA lines, as in HMMER
    rc = subprocess.run([sys.executable, exe, "--cut_ga", str(tmp_path / "g.hmm"), str(tmp_path / "t.fasta")], capture_output=True, text=True)
    assert rc.returncode == 1 and "GA bit thresholds unavailable" in rc.stderr


def test_front_end_argument_errors_without_a_gpu(tmp_path, capsys):
    """Usage and option errors are reported with exit status 1 before any device is touched (lib/aln.py:669-672 then logs
    'Error in hmmsearch command'); a missing GPU is an error as well, never a CPU fallback."""
    from phyloaln_b200 import frontend
    assert frontend.hmmsearch_main([]) == 1
    assert frontend.hmmsearch_main(["--cpu"]) == 1
    assert frontend.hmmsearch_main(["--cut_ga", "-E", "1", "a.hmm", "b.fa"]) == 1
    assert frontend.hmmsearch_main(["--frobnicate", "a.hmm", "b.fa"]) == 1
    assert frontend.hmmsearch_main(["-o", str(tmp_path / "o.txt"), str(tmp_path / "missing.hmm"), str(tmp_path / "missing.fa")]) == 1
    err = capsys.readouterr().err
    assert "Usage" in err and "cut" in err and not (tmp_path / "o.txt").exists()


def _gappy_alignment(path):
    ref = [l for l in open(os.path.join(HERE, "golden", "config1", "ref", "OG0003709.ref.fas")).read().split(">") if l]
    rows = {}
    with open(path, "w") as fh:
        for r in ref:
            n, s = r.split("\n", 1)
            s = s.replace("\n", "")
            s = s[:20] + ("--" if n.startswith("D") else "QW") + s[20:]     # an insert column pair (one sequence only)
            rows[n] = s
            fh.write(f">{n}\n{s}\n")
    return rows


def test_hmmbuild_front_end_files(tmp_path, oracle):
    """bin/hmmbuild's outputs as the reference consumes them: the .hmm loads, and the -O Stockholm file's RF line selects the
    match columns the way lib/aln.py:767-782 does (RF != '.', upper case, '.'/'~' -> '-')."""
    from phyloaln_b200 import frontend, hmmio
    rows = _gappy_alignment(tmp_path / "a.fas")
    rc = frontend.hmmbuild_main(["-O", str(tmp_path / "a.sto"), "-n", "OGx", "--cpu", "2", "-o", str(tmp_path / "a.log"),
                                 str(tmp_path / "a.hmm"), str(tmp_path / "a.fas")],
                                score_fn=lambda hmm, seqs: oracle.Profile.from_hmm(hmm).score_all(seqs))
    assert rc == 0
    hmm = hmmio.read_hmm(str(tmp_path / "a.hmm"))
    assert hmm.name == "OGx" and hmm.M == 128 and set(hmm.stats) == {"MSV", "VITERBI", "FORWARD"}
    sto = frontend.read_alignment(str(tmp_path / "a.sto"))
    rf = [l.split()[2] for l in open(tmp_path / "a.sto") if l.startswith("#=GC RF")][0]
    assert rf.count("x") == hmm.M and len(rf) == 130 and rf[20:22] == ".."
    for sid, s in rows.items():
        picked = "".join(ch for ch, m in zip(sto[sid], rf) if m != ".").upper().replace(".", "-").replace("~", "-")
        assert picked == s[:20] + s[22:]                      # = the original reference columns
        assert sto[sid][20:22] == ("qw" if not sid.startswith("D") else "..")
    assert frontend.hmmbuild_main(["only_one_arg"]) == 1 and frontend.guess_alphabet({"a": "ACGTNACGT-"}) == "dna"


@pytest.mark.gpu
def test_hmmbuild_then_hmmsearch_front_ends_on_the_gpu(tmp_path):
    """The two executables chained as the unmodified PhyloAln would spawn them: profile from an alignment, then a search of
    translated reads; reads derived from the alignment must be found."""
    import subprocess
    from oracle import oracle as O
    from phyloaln_b200 import synth
    rows = _gappy_alignment(tmp_path / "a.fas")
    bindir = os.path.join(os.path.dirname(HERE), "phyloaln_b200", "bin")
    rc = subprocess.run([sys.executable, os.path.join(bindir, "hmmbuild"), "-O", str(tmp_path / "a.sto"), "-n", "OGx", "--cpu", "1",
                         str(tmp_path / "a.hmm"), str(tmp_path / "a.fas")], capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr
    prot = next(iter(rows.values())).replace("-", "")
    rng = np.random.default_rng(2)
    with open(tmp_path / "t.fasta", "w") as fh:
        for i in range(40):
            a = int(rng.integers(0, len(prot) - 45))
            nt = synth.back_translate(prot[a:a + 45])
            for suffix, p in O.six_frames(nt, 1):
                fh.write(f">r{i}_PhAlTag_S_fastx1{suffix}\n{p}\n")
    rc = subprocess.run([sys.executable, os.path.join(bindir, "hmmsearch"), "-o", str(tmp_path / "o.txt"), "--tblout", str(tmp_path / "o.tbl"),
                         "--cpu", "1", "-E", "1e-3", str(tmp_path / "a.hmm"), str(tmp_path / "t.fasta")], capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr
    sys.path.insert(0, SHIM)
    try:
        from Bio import SearchIO
        hits = [f.hit_id for q in SearchIO.parse(str(tmp_path / "o.txt"), "hmmer3-text") for hit in q for hsp in hit for f in hsp]
    finally:
        sys.path.remove(SHIM)
    assert len(hits) >= 35 and all(h.endswith("_pos1") for h in hits)
