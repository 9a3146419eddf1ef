"""Reference-column definition (/root/reference/lib/aln.py:745-792): phyloaln_b200.refcols against outputs of
the reference's own statements (tests/golden/make_refcols_golden.py runs the block of lib/aln.py it locates at generation
time) -- files byte for byte, and an error wherever the reference's `except:` would report "fail to convert"."""
import gzip
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    with gzip.open(os.path.join(HERE, "golden", "refcols_golden.json.gz"), "rt") as fh:
        return json.load(fh)["cases"]


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_reference_columns_equal_the_reference(case, tmp_path):
    from phyloaln_b200 import refcols
    os.makedirs(tmp_path / "ref_hmm")
    (tmp_path / "ref_hmm" / "g.sto").write_text(case["sto"])
    codon = None
    if "codon" in case:
        codon = str(tmp_path / "codon.fas")
        with open(codon, "w") as fh:
            fh.write(case["codon"])
    rna = case.get("hmmtool", "hmmbuild").startswith("cm") and not case.get("infernal_no_replace", False)
    want = case["want"]
    if "error" in want:
        with pytest.raises(ValueError):      # StockholmError: the whole step fails loudly, as (1, group) does in the reference
            refcols.reference_columns_step(str(tmp_path), ["g"], {"g": codon} if codon else None, rna_to_dna=rna)
        return
    n = refcols.reference_columns_step(str(tmp_path), ["g"], {"g": codon} if codon else None, rna_to_dna=rna)
    assert (tmp_path / "map" / "g.ref.fas").read_text() == want["ref_fas"]
    rows = want["ref_fas"].split("\n")
    assert n["g"] == len(rows[1])
    if "codon_ref_fas" in want:
        assert (tmp_path / "map" / "g.codon.ref.fas").read_text() == want["codon_ref_fas"]
    else:
        assert not (tmp_path / "map" / "g.codon.ref.fas").exists()


def test_existing_ref_fas_is_copied(tmp_path):
    """lib/aln.py:745-749: a `ref_hmm/{g}.ref.fas` (HMM-database mode) wins over the .sto and is copied verbatim."""
    from phyloaln_b200 import refcols
    os.makedirs(tmp_path / "ref_hmm")
    text = ">a desc\nAC-GT\nac\n>b\n\nTT\n"
    (tmp_path / "ref_hmm" / "g.ref.fas").write_text(text)
    (tmp_path / "ref_hmm" / "g.sto").write_text("garbage")
    refcols.reference_columns_step(str(tmp_path), ["g"])
    assert (tmp_path / "map" / "g.ref.fas").read_text() == text


def test_columns_of_the_lite_hmmbuild_match_its_profile(tmp_path):
    """The chain the drop-in relies on: `bin/hmmbuild -O` (lite) writes the .sto with the match columns it gave the profile,
    and the reference columns defined from that .sto are exactly the profile's M nodes, residue for residue."""
    from phyloaln_b200 import frontend, hmmbuild_lite as hb, refcols
    os.makedirs(tmp_path / "ref_hmm")
    aln = tmp_path / "in.fas"
    aln.write_text(">s1\nMKV-LAAGIT\n>s2\nMKVQLA-GIT\n>s3\nM-V-LAAGLT\n>s4\nMRV-LSAGIT\n")
    rows = frontend.read_alignment(str(aln))
    hmm, mcols = hb.build_from_alignment("g", rows, "amino")       # the model construction of hmmbuild_main, without calibration
    frontend.write_stockholm(str(tmp_path / "ref_hmm" / "g.sto"), "g", rows, mcols)
    n = refcols.reference_columns_step(str(tmp_path), ["g"])
    assert n["g"] == hmm.M == len(mcols)
    ref, _ = refcols.reference_columns(str(tmp_path / "ref_hmm" / "g.sto"))
    assert list(ref) == ["s1", "s2", "s3", "s4"]
    for sid, row in rows.items():
        assert ref[sid] == "".join(row[c] for c in mcols).upper()
