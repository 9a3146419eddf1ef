"""merge_hits() (the contig fold of /root/reference/lib/assemble.py:889-954) through pa_merge_hists: 242 chains folded by the
reference's own Python (tests/golden/make_merge_golden.py) -- codon and nucleotide mode, fill_gap '-' / 'N', --ref_split_len
interval merging, a 'too much' item -- must come back identical, dictionary insertion order included."""
import gzip
import json
import os

import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "merge_golden.json.gz")


def _item(d, codon):
    if codon:
        comp = {int(c): {k + 1: {nt: n for nt, n in lst} for k, lst in enumerate(ks)} for c, ks in d["seq_comp"]}
    else:
        comp = {int(c): {nt: n for nt, n in lst} for c, lst in d["seq_comp"]}
    return {"qstart": d["qstart"], "qend": d["qend"], "seq_comp": comp, "seqids": {k: [list(iv) for iv in v] for k, v in d["seqids"]},
            "count": {k: v for k, v in d["count"]}}


def _ordered(item, codon):
    if codon:
        comp = [[c, [[[nt, n] for nt, n in d[k].items()] for k in (1, 2, 3)]] for c, d in item["seq_comp"].items()]
    else:
        comp = [[c, [[nt, n] for nt, n in d.items()]] for c, d in item["seq_comp"].items()]
    return {"qstart": item["qstart"], "qend": item["qend"], "seq_comp": comp, "seqids": [[k, v] for k, v in item["seqids"].items()],
            "count": [[k, v] for k, v in item["count"].items()]}


def test_merge_hits_chains_match_reference_python(gpu_ctx_factory):
    from phyloaln_b200 import mapping
    gold = json.load(gzip.open(GOLD, "rt"))["chains"]
    ctx = gpu_ctx_factory()
    groups = {}
    for c in gold:
        groups.setdefault((c["mol_type"], c["split_len"], c["fill_gap"]), []).append(c)
    assert len(groups) >= 6 and len(gold) >= 200
    for (mol, split, fill), cases in groups.items():
        codon = mol != "dna"
        got = mapping.merge_hits_chains(ctx, [[_item(i, codon) for i in c["items"]] for c in cases], mol_type=mol, split_len=split, fill_gap=fill)
        for c, g in zip(cases, got):
            assert _ordered(g, codon) == c["result"], (c["group"], c["species"], mol, split, fill, len(c["items"]))
    ctx.close()
