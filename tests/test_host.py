"""CPU tests of the host logic and of the C-ABI surface (no GPU compute)."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_every_declared_symbol():
    from phyloaln_b200 import capi
    L = capi.load_library()
    header = open(os.path.join(ROOT, "include", "phyloaln_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pa_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 19
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/phyloaln_b200.h but not exported"
    assert set(capi.EXPORTS) <= set(declared)
    assert L.pa_abi_version() == 3


def test_no_gpu_is_a_loud_error():
    from phyloaln_b200 import capi
    try:
        ctx = capi.Context(0)
    except capi.PaError as e:       # build container: no device -> must refuse, never fall back
        assert "CUDA" in str(e) or "device" in str(e)
        return
    ctx.close()                     # on a GPU box the context simply works


def test_struct_layouts_match_header():
    from phyloaln_b200 import capi
    import ctypes
    assert ctypes.sizeof(capi.PaHit) == 120 and ctypes.sizeof(capi.PaMapRec) == 64 and ctypes.sizeof(capi.PaMapOpts) == 32 and ctypes.sizeof(capi.PaAccum) == 56 and ctypes.sizeof(capi.PaOpts) == 40
    assert capi.HIT_DTYPE.itemsize == 120


def test_parse_hmmsearch_parameters():
    from phyloaln_b200.search import SearchParameterError, parse_hmmsearch_parameters as p
    o = p([" -E", "1e-5", " --domE", "0.1", " --nobias", " --cpu", "4"])
    assert (o.E, o.domE, o.nobias, o.max, o.T) == (1e-5, 0.1, True, False, None) and o.ignored == ["--cpu"]
    o = p(["-T 25 --domT 20 -Z 1000 --domZ 10 --F1 0.1 --F2 0.01 --F3 1e-4 --max --nonull2"])
    assert (o.T, o.domT, o.Z, o.domZ, o.F1, o.F2, o.F3, o.max, o.nonull2) == (25, 20, 1000, 10, 0.1, 0.01, 1e-4, True, True)
    assert p([]).E == 10.0 and p(None).domE == 10.0
    assert (p(["--cut_ga"]).cut, p([" --cut_tc", "--cpu", "2"]).cut, p(["--cut_nc --nobias"]).cut, p([]).cut) == ("GA", "TC", "NC", None)
    for bad in (["--cut_ga", "-E", "1"], ["--cut_ga --cut_tc"], ["--domT", "3", "--cut_nc"], ["-E"], ["-E", "abc"], ["--bogus"]):
        with pytest.raises(SearchParameterError):
            p(bad)


def test_model_cutoffs_replace_the_evalue_thresholds(tmp_path):
    """--cut_ga / --cut_tc / --cut_nc (HMMER's p7_pli_NewModelThresholds): every profile reports with ITS OWN per-sequence and
    per-domain bit-score thresholds, read from the GA / TC / NC lines of its file -- the same selection as -T / --domT with
    those values, profile by profile; a profile without the lines is an error."""
    from phyloaln_b200 import capi, hmmio, search, synth
    rng = np.random.default_rng(12)
    hmms = []
    for i, cut in enumerate([{"GA": (25.0, 20.5), "TC": (30.0, 28.0)}, {"GA": (12.25, 3.0), "TC": (40.0, 1.0)}, {"GA": (60.0, 0.0), "TC": (0.0, 50.0)}]):
        h = synth.random_hmm(f"c{i}", 20 + i, rng)
        h.stats = {"MSV": (-8.0, 0.71), "VITERBI": (-9.0, 0.71), "FORWARD": (-4.0, 0.71)}
        h.cutoffs = cut
        path = str(tmp_path / f"c{i}.hmm")
        hmmio.write_hmm(h, path)
        for rd in (hmmio.read_hmm(path), hmmio.read_hmm_py(path)):      # native reader + header scan, and the Python statement
            assert rd.cutoffs == cut
        hmms.append(hmmio.read_hmm(path))
    n = 600
    hits = np.zeros(n, dtype=capi.HIT_DTYPE)
    hits["hmm"] = rng.integers(0, 3, n)
    hits["target"] = rng.integers(0, 150, n)
    order = np.lexsort((hits["target"], hits["hmm"]))
    hits = hits[order]
    key = hits["hmm"].astype(np.int64) * 1000 + hits["target"]
    first = np.r_[True, key[1:] != key[:-1]]
    grp = np.cumsum(first) - 1
    hits["dom_idx"] = np.arange(n) - np.flatnonzero(first)[grp]
    seq_bits = rng.uniform(0, 70, grp.max() + 1).astype(np.float32)
    seq_bits[:6] = [25.0, 12.25, 60.0, 30.0, 24.99, 12.24]            # values sitting on the thresholds
    hits["seq_bits"] = seq_bits[grp]
    hits["seq_lnP"] = -hits["seq_bits"]
    hits["dom_bits"] = rng.uniform(-5, 60, n).astype(np.float32)
    hits["dom_bits"][:4] = [20.5, 3.0, 0.0, 28.0]
    hits["dom_lnP"] = -hits["dom_bits"]
    name = lambda t: f"t{t:04d}"
    for flag, tag in (("--cut_ga", "GA"), ("--cut_tc", "TC")):
        o = search.parse_hmmsearch_parameters([flag])
        search.resolve_cutoffs(o, hmms)
        _, got = search.report_index(hits, name, 150, o, 3)
        assert sum(len(v[0]) for v in got.values()) > 20
        for h in range(3):
            a, b = hmms[h].cutoffs[tag]
            ot = search.parse_hmmsearch_parameters(["-T", repr(a), "--domT", repr(b)])
            _, want = search.report_index(hits[hits["hmm"] == h], name, 150, ot, 3)
            base = np.flatnonzero(hits["hmm"] == h)
            if h in want:
                assert np.array_equal(got[h][0], base[want[h][0]]) and got[h][1] == want[h][1]
            else:
                assert h not in got
    o = search.parse_hmmsearch_parameters(["--cut_nc"])
    with pytest.raises(search.SearchParameterError, match="NC bit thresholds unavailable on model c0"):
        search.resolve_cutoffs(o, hmms)
    with pytest.raises(search.SearchParameterError, match="not resolved"):
        search.report_index(hits, name, 150, o, 3)


def test_frame_naming_and_rows():
    from phyloaln_b200 import search
    assert [search.frame_suffix(f, 6, True) for f in range(6)] == [("_pos1", "+", 0), ("_pos2", "+", 1), ("_pos3", "+", 2),
                                                                    ("_rev_pos1", "rev", 0), ("_rev_pos2", "rev", 1), ("_rev_pos3", "rev", 2)]
    assert search.frame_suffix(0, 1, True) == ("_pos1", "+", 0)                     # --codon
    assert [search.frame_suffix(f, 3, True) for f in range(3)] == [("_pos1", "+", 0), ("_pos2", "+", 1), ("_pos3", "+", 2)]
    assert [search.frame_suffix(f, 2, False) for f in range(2)] == [("", "+", None), ("_rev", "rev", None)]
    rows = [{"target": 9, "hmmfrom": 2, "hmmto": 9, "alifrom": 1, "alito": 9, "model": "ktayi.akq", "aligned": "KTAYIwAKQ"}]
    assert search.hit_rows("g", rows, ["r0", "r1"], 6, True) == [("g", "r1", 2, 9, 1, 9, "rev", 0, "KTAYI-AKQ", "KTAYIWAKQ")]
    assert search.hit_rows("g", [dict(rows[0], target=1)], ["r0"], 2, False)[0][6:8] == ("rev", None)
    assert search.shard_bounds(10, 4, 0) == (0, 3) and search.shard_bounds(10, 4, 3) == (9, 10) and search.shard_bounds(2, 4, 3) == (2, 2)


def _fake_hits(oracle):
    """Oracle hits of two profiles recast into the product's hit record layout."""
    from phyloaln_b200 import capi
    from tests import helpers
    rng = np.random.default_rng(21)
    hmms = [helpers.calibrated_hmm(f"rp{i}", M, seed=70 + i, bg_mix=0.1) for i, M in enumerate([60, 110])]
    seqs = helpers.protein_targets(hmms, 600, rng, Lmin=30, Lmax=60, plant_every=4)
    names = [f"read{i // 6}_pos{i % 6}" for i in range(len(seqs))]
    recs, model, aligned, per = [], [], [], []
    for hi, h in enumerate(hmms):
        oh, _ = oracle.Profile.from_hmm(h).search(seqs, nthreads=4, want_trace=False)
        per.append(oh)
        for d in oh:
            r = np.zeros(1, dtype=capi.HIT_DTYPE)[0]
            r["hmm"] = hi
            for f in ("target", "ndom_in_seq", "dom_idx", "ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto", "dom_bits", "dom_lnP",
                      "seq_bits", "seq_lnP", "seqbias", "dombias"):
                r[f] = d[f]
            recs.append(r)
            model.append(d["model"])
            aligned.append(d["aligned"])
    return hmms, seqs, names, np.array(recs, dtype=capi.HIT_DTYPE), model, aligned, per


def test_report_matches_oracle_report(oracle):
    from oracle import report as orep
    from phyloaln_b200 import search
    hmms, seqs, names, hits, model, aligned, per = _fake_hits(oracle)
    Z = 5.0e6   # large Z so that the E-value thresholds actually cut
    perm = np.random.default_rng(1).permutation(len(hits))
    rep = search.report(hits[perm], [model[i] for i in perm], [aligned[i] for i in perm], lambda t: names[t], Z,
                        search.SearchOptions(), len(hmms))
    kept = 0
    for hi in range(len(hmms)):
        want = orep.apply_report(per[hi], names, Z)
        got = rep[hi]
        assert [(w["target"], w["dom_idx"]) for w in want] == [(g["target"], g["dom_idx"]) for g in got]
        assert [w["model"] for w in want] == [g["model"] for g in got]
        kept += len(got)
        assert len(got) < len(per[hi])          # thresholds removed something
    assert kept > 20
    # explicit score thresholds
    o = search.SearchOptions(T=30.0, domT=25.0)
    rep = search.report(hits, model, aligned, lambda t: names[t], Z, o, len(hmms))
    for hi in range(len(hmms)):
        assert all(r["seq_bits"] >= 30.0 and r["dom_bits"] >= 25.0 for r in rep[hi])


def test_write_step_outputs_and_resume_markers(tmp_path, oracle):
    from phyloaln_b200 import search
    rows = [{"target": 7, "hmmfrom": 2, "hmmto": 9, "alifrom": 1, "alito": 9, "model": "KTAYI.AKQ", "aligned": "KTAYIwAKQ",
             "name": "r1_pos2", "seq_evalue": 1e-9, "seq_bits": 40.0, "seq_bias": 0.1, "dom_ievalue": 1e-9, "dom_bits": 39.0,
             "dom_bias": 0.1, "ndom": 1},
            {"target": 7, "hmmfrom": 20, "hmmto": 22, "alifrom": 12, "alito": 14, "model": "AAA", "aligned": "AAA",
             "name": "r1_pos2", "seq_evalue": 1e-9, "seq_bits": 40.0, "seq_bias": 0.1, "dom_ievalue": 1e-3, "dom_bits": 9.0,
             "dom_bias": 0.0, "ndom": 1}]
    search.write_step_outputs(str(tmp_path), "g", rows, ["r0", "r1"], lambda i: ["ACGT", "GGCC"][i], 6, True, ["log line"])
    assert open(tmp_path / "map" / "g.hit_info.tsv").read() == "g\tr1\t2\t9\t1\t9\t+\t1\tKTAYI-AKQ\tKTAYIWAKQ\ng\tr1\t20\t22\t12\t14\t+\t1\tAAA\tAAA\n"
    assert open(tmp_path / "map" / "g.targets.fas").read() == ">r1\nGGCC\n"
    assert os.path.isfile(tmp_path / "ok" / "hmmsearch_g.ok") and os.path.isfile(tmp_path / "ok" / "extract_reads_g.ok")
    search.write_tblout(str(tmp_path / "g.tbl"), "g", rows)
    assert "r1_pos2" in open(tmp_path / "g.tbl").read()


def test_read_raw_fasta_dictionary_semantics(tmp_path):
    from phyloaln_b200 import dropin
    p = tmp_path / "all.raw.fasta"
    p.write_text(">a_PhAlTag_S_fastx1\nACGT\n>b_PhAlTag_S_fastx1\nGG\n>a_PhAlTag_S_fastx1\nTTTTT\n")
    ids, nt, off = dropin.read_raw_fasta(str(p))
    assert ids == ["a_PhAlTag_S_fastx1", "b_PhAlTag_S_fastx1"] and nt.tobytes() == b"TTTTTGG" and list(off) == [0, 5, 7]


def test_hmmbuild_lite_and_file_roundtrip(tmp_path):
    from phyloaln_b200 import hmmbuild_lite as hb, hmmio
    aln = {"A.1": "MKTAYIAKQRQI--SF", "B.1": "MKTAYIAKQRQIGGSF", "C.1": "MKSAYIGKQRQL--SF", "OUT.1": "MRSAF-GKNRQL--AF"}
    hmm, mcols = hb.build_from_alignment("t", aln, "amino")
    assert hmm.M == 14 and mcols == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 14, 15]
    p = np.exp(-hmm.mat[1:])
    assert np.allclose(p.sum(1), 1, atol=1e-6) and np.allclose(np.exp(-hmm.t[:, 0:3]).sum(1), 1, atol=1e-6)
    assert hmm.consensus().upper().startswith("MK") and math.isinf(hmm.t[0][6]) and hmm.t[hmm.M][5] < 1e-9
    with pytest.raises(hmmio.HMMFormatError):
        hmm.evparams()
    hmm.stats = {"MSV": (-9.5, 0.71), "VITERBI": (-10.2, 0.71), "FORWARD": (-4.1, 0.71)}
    hmmio.write_hmm(hmm, str(tmp_path / "t.hmm"))
    h2 = hmmio.read_hmm(str(tmp_path / "t.hmm"))
    assert h2.M == hmm.M and h2.cons == hmm.cons and h2.map == hmm.map and list(h2.evparams()) == list(hmm.evparams())
    assert np.allclose(h2.mat[1:], hmm.mat[1:], atol=6e-6) and np.array_equal(np.isinf(h2.t), np.isinf(hmm.t))
    with pytest.raises(hmmio.HMMFormatError):
        (tmp_path / "bad.hmm").write_text("not an hmm\n")
        hmmio.read_hmm(str(tmp_path / "bad.hmm"))
    dna, _ = hb.build_from_alignment("d", {"a": "ACGTACGTNN", "b": "ACGTACGTAC"}, "dna")
    assert dna.K == 4 and dna.M == 10


def test_composition_table():
    from phyloaln_b200 import mapping
    tab = mapping.composition_table({1: {"M": 3}, 2: {"K": 2, "-": 1}, 3: {"*": 1, "X": 2}})
    assert tab.shape == (3, 32) and tab[0, 12] == 3 and tab[1, 26] == 1 and tab[2, 27] == 1 and tab[2, 23] == 2


def test_ref_score_tables_match_reference_get_ref_score():
    """mapping.ref_score_tables == get_ref_score() (lib/assemble.py:678-764): against the committed vectors the reference
    produced for the five config-1 alignments, and - when the reference tree is present - against the reference itself on
    gappy random alignments (valid ranges, several outgroups, explicit ingroup, missing species)."""
    import gzip, json, sys
    from phyloaln_b200 import mapping
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(gzip.open(os.path.join(here, "golden", "mapping_ext_golden.json.gz"), "rt"))

    def read_fasta(path):
        seqs, name = {}, None
        for line in open(path):
            line = line.rstrip("\n")
            if line.startswith(">"):
                name = line[1:].split()[0]
                seqs[name] = ""
            elif name:
                seqs[name] += line
        return seqs
    for g, G in gold["groups"].items():
        ref = read_fasta(os.path.join(here, "golden", "config1", "ref", g + ".ref.fas"))
        comp, names, tab = mapping.ref_score_tables(ref, outgroup=["SLEBA"], sep=".", unknow="X")
        assert (comp == mapping.composition_table({int(k): v for k, v in G["site_composition"].items()})).all()
        assert names == list(G["outgroup_score"].keys())
        want = mapping.outgroup_table({s: {int(c): v for c, v in d.items()} for s, d in G["outgroup_score"].items()}, comp.shape[0])
        assert np.array_equal(np.isnan(tab), np.isnan(want)) and np.array_equal(tab[~np.isnan(tab)], want[~np.isnan(want)])
    if not os.path.isdir("/root/reference/lib"):
        return
    sys.path.insert(0, os.path.join(here, "bio_shim"))
    sys.path.insert(0, "/root/reference/lib")
    try:
        import assemble as asm
    finally:
        sys.path.remove("/root/reference/lib")
        sys.path.remove(os.path.join(here, "bio_shim"))
    rng = np.random.default_rng(3)
    n = 0
    for t in range(30):
        ncols, nseq = int(rng.integers(20, 120)), int(rng.integers(3, 9))
        ref = {}
        for s in range(nseq):
            row = list("ACDEFGHIKLMNPQRSTVWY"[int(x)] for x in rng.integers(0, 20, ncols))
            for _ in range(int(rng.integers(0, 6))):          # gap runs of assorted lengths, some X
                a, ln = int(rng.integers(0, ncols)), int(rng.integers(1, 12))
                row[a:a + ln] = ("-" if rng.random() < 0.8 else "X") * len(row[a:a + ln])
            ref[f"SP{s}.{t}"] = "".join(row)
        for kw in (dict(outgroup=["SP0"]), dict(outgroup=[]), dict(outgroup=["SP0", "SP1"], ingroup=["SP2", "SP3", "SPX"]),
                   dict(outgroup=["NOPE", "SP1"])):
            try:
                sc, osc = asm.get_ref_score(dict(ref), sep=".", unknow="X", **kw)
            except (ZeroDivisionError, ValueError):
                with pytest.raises((ZeroDivisionError, ValueError)):
                    mapping.ref_score_tables(ref, sep=".", unknow="X", **kw)
                continue
            comp, names, tab = mapping.ref_score_tables(ref, sep=".", unknow="X", **kw)
            assert (comp == mapping.composition_table(sc)).all() and names == list(osc.keys())
            want = mapping.outgroup_table(osc, ncols)
            assert np.array_equal(np.isnan(tab), np.isnan(want)) and np.array_equal(tab[~np.isnan(tab)], want[~np.isnan(want)]), (t, kw)
            n += 1
    assert n > 60


def test_native_hmm_reader_equals_python_statement(tmp_path):
    """pa_hmm_file_parse (hmmio.read_hmm) against hmmio.read_hmm_py: config-1 profiles, a written DNA profile without
    STATS, numbers in exponent notation, and malformed files (same exception from both)."""
    import glob
    from phyloaln_b200 import hmmio, synth
    here = os.path.dirname(os.path.abspath(__file__))
    files = sorted(glob.glob(os.path.join(here, "golden", "config1", "ref_hmm", "*.hmm")))
    dna = synth.random_hmm("dnaprof", 77, np.random.default_rng(4), abc="dna")
    dna.stats = {}
    hmmio.write_hmm(dna, str(tmp_path / "dna.hmm"))
    files.append(str(tmp_path / "dna.hmm"))
    src = open(files[0]).read()
    tokv = [l for l in src.splitlines() if l.split() and l.split()[0] == "1"][0].split()[1]   # first match emission of node 1
    assert src.count(tokv) >= 1
    (tmp_path / "exp.hmm").write_text(src.replace(tokv, tokv + "e+00"))
    files.append(str(tmp_path / "exp.hmm"))
    for f in files:
        a, b = hmmio.read_hmm(f), hmmio.read_hmm_py(f)
        assert (a.name, a.abc, a.M, a.cons, a.rf, a.map, a.nseq, a.effn) == (b.name, b.abc, b.M, b.cons, b.rf, b.map, b.nseq, b.effn), f
        assert np.array_equal(a.mat, b.mat) and np.array_equal(a.ins, b.ins) and np.array_equal(a.t, b.t), f
        assert (a.compo is None) == (b.compo is None) and (a.compo is None or np.array_equal(a.compo, b.compo))
        assert set(a.stats) == set(b.stats)
        for k in a.stats:
            assert a.stats[k] == pytest.approx(b.stats[k], rel=1e-6)
    good = open(files[0]).read().splitlines()
    for name, lines in (("trunc", good[:len(good) // 2]), ("nohdr", good[1:]), ("badnum", [l.replace(tokv, tokv[:2] + "x" + tokv[3:]) for l in good]),
                        ("noleng", [l for l in good if not l.startswith("LENG")])):
        (tmp_path / (name + ".hmm")).write_text("\n".join(lines) + "\n")
        with pytest.raises((hmmio.HMMFormatError, ValueError, IndexError)) as e1:
            hmmio.read_hmm_py(str(tmp_path / (name + ".hmm")))
        with pytest.raises(type(e1.value)):
            hmmio.read_hmm(str(tmp_path / (name + ".hmm")))


def test_msv_work_item_order_is_a_bijection_with_l2_super_tiles():
    """pa_msv_item_decode (the function the MSV kernels run): every (profile slot, target tile) pair exactly once, tiles of a
    super-tile contiguous in item order, profile-major inside a super-tile, plain profile-major order when one super-tile
    covers the chunk."""
    import ctypes as C
    from phyloaln_b200 import capi
    L = capi.load_library()
    L.pa_msv_item_decode.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    slot, tile = C.c_uint32(), C.c_uint32()
    for nlist, ntiles, T in ((7, 10, 3), (1, 1, 1), (5, 9, 9), (5, 9, 100), (3, 12, 4), (466, 37, 5), (2, 1536, 461)):
        seen, order = set(), []
        for item in range(nlist * ntiles):
            assert L.pa_msv_item_decode(item, nlist, ntiles, T, C.byref(slot), C.byref(tile)) == 0
            assert slot.value < nlist and tile.value < ntiles
            seen.add((slot.value, tile.value))
            order.append((slot.value, tile.value))
        assert len(seen) == nlist * ntiles
        supers = [t // T for _, t in order]
        assert supers == sorted(supers)                               # super-tile-major
        for s in set(supers):                                         # profile-major inside a super-tile
            sub = [o for o, sp in zip(order, supers) if sp == s]
            assert sub == sorted(sub)
        if T >= ntiles:
            assert order == [(p, t) for p in range(nlist) for t in range(ntiles)]
    assert L.pa_msv_item_decode(70, 7, 10, 3, C.byref(slot), C.byref(tile)) == -2
