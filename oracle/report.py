"""ORACLE (test infrastructure): hmmsearch report thresholds + order + hit_info rows.

Independent restatement of SURVEY.md A.9/A.10 and of extract_reads() (reference
lib/aln.py:544-558) used to turn oracle hits into golden hit_info.tsv rows.  Never imported by
the product (phyloaln_b200/search.py has its own implementation)."""
import math


def apply_report(hits, names, Z, E=10.0, domE=10.0):
    """hits: list of oracle hit dicts of ONE profile; names[target] -> target name.
    Returns reported domains in hmmsearch order."""
    by_t = {}
    for h in hits:
        by_t.setdefault(h["target"], []).append(h)
    seqs = [t for t, hs in by_t.items() if math.exp(hs[0]["seq_lnP"]) * Z <= E]
    domZ = float(len(seqs))
    seqs.sort(key=lambda t: (by_t[t][0]["seq_lnP"], names[t].encode()))
    out = []
    for t in seqs:
        for h in sorted(by_t[t], key=lambda d: d["dom_idx"]):
            if math.exp(h["dom_lnP"]) * domZ <= domE:
                out.append(h)
    return out


def rows_from_hits(query, hits, names):
    """10-column rows as lib/aln.py:558 writes them (suffix stripping order of :546-551)."""
    rows = []
    for h in hits:
        hit_id = names[h["target"]]
        strand, start_pos = "+", None
        if hit_id.endswith(("_pos1", "_pos2", "_pos3")):
            start_pos = int(hit_id[-1]) - 1
            hit_id = hit_id[:-5]
        if hit_id.endswith("_rev"):
            hit_id = hit_id[:-4]
            strand = "rev"
        rows.append((query, hit_id, h["hmmfrom"], h["hmmto"], h["iali"], h["jali"], strand, start_pos,
                     h["model"].upper().replace(".", "-"), h["aligned"].upper().replace(".", "-")))
    return rows
