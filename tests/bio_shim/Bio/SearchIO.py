"""Test-only stand-in for Bio.SearchIO.parse(handle, 'hmmer3-text') (Biopython is not in this image).

It reads the hmmsearch text report the way Biopython's Hmmer3TextParser does -- query line, '>> hit' sections, the
per-domain table (coordinates become 0-based half-open: start = from - 1), and the '== domain' alignment blocks
(model line / match line / target line / PP line, concatenated over blocks) -- and exposes the attributes the
reference's extract_reads() uses (lib/aln.py:511-545): iteration qresult -> hit -> hsp -> fragment with query_id,
hit_id, query_start/end, hit_start/end, query.seq, hit.seq.
"""
import re


class _Seq:
    def __init__(self, seq):
        self.seq = seq


class _Frag:
    def __init__(self, query_id, hit_id):
        self.query_id, self.hit_id = query_id, hit_id
        self.query_start = self.query_end = self.hit_start = self.hit_end = None
        self.query, self.hit = _Seq(""), _Seq("")
        self.hit_strand = 0
        self.aln_annotation = {}


class _HSP(list):
    pass


class _Hit(list):
    def __init__(self, hit_id):
        super().__init__()
        self.id = hit_id


class _QResult(list):
    def __init__(self, qid):
        super().__init__()
        self.id = qid


_DOM = re.compile(r"^\s*(\d+)\s+([!?])\s+(-?[\d.]+)\s+(-?[\d.]+)\s+(\S+)\s+(\S+)\s+(\d+)\s+(\d+)\s+(\S\S)\s+(\d+)\s+(\d+)\s+(\S\S)\s+(\d+)\s+(\d+)\s+(\S\S)\s+([\d.]+)")
_ALN = re.compile(r"^\s+(\S+)\s+(\d+|-)\s+(\S+)\s+(\d+|-)\s*$")


def parse(handle, fmt):
    if fmt != "hmmer3-text":
        raise ValueError("the test shim only parses hmmer3-text")
    fh = open(handle) if isinstance(handle, str) else handle
    lines = [l.rstrip("\n") for l in fh]
    if isinstance(handle, str):
        fh.close()
    results, q, i = [], None, 0
    while i < len(lines):
        line = lines[i]
        if line.startswith("Query:"):
            q = _QResult(line.split()[1])
            results.append(q)
        elif line.startswith(">> ") and q is not None:
            hit = _Hit(line[3:].split()[0])
            q.append(hit)
            i += 1
            doms = []
            while i < len(lines) and not lines[i].lstrip().startswith("Alignments for each domain") and not lines[i].startswith(">> "):
                m = _DOM.match(lines[i])
                if m:
                    doms.append(m)
                i += 1
            for m in doms:
                f = _Frag(q.id, hit.id)
                f.query_start, f.query_end = int(m.group(7)) - 1, int(m.group(8))
                f.hit_start, f.hit_end = int(m.group(10)) - 1, int(m.group(11))
                hsp = _HSP([f])
                hit.append(hsp)
            # alignment blocks
            d = -1
            while i < len(lines) and not lines[i].startswith(">> ") and not lines[i].startswith("Internal pipeline") and not lines[i].startswith("//"):
                s = lines[i]
                if s.lstrip().startswith("== domain"):
                    d = int(s.split()[2]) - 1
                    i += 1
                    continue
                m = _ALN.match(s) if d >= 0 else None
                if m and m.group(1) == q.id and not s.rstrip().endswith((" PP", " RF", " CS")):
                    frag = hit[d][0]
                    frag.query.seq += m.group(3)
                    col = s.index(m.group(3), s.index(m.group(2), len(s) - len(s.lstrip()) + len(m.group(1))) + len(m.group(2)))
                    i += 2                                  # skip the match line
                    t = _ALN.match(lines[i])
                    assert t is not None, lines[i]
                    frag.hit.seq += t.group(3)
                    if i + 1 < len(lines) and lines[i + 1].rstrip().endswith(" PP"):
                        frag.aln_annotation["PP"] = frag.aln_annotation.get("PP", "") + lines[i + 1][col:col + len(m.group(3))]
                        i += 1
                i += 1
            continue
        i += 1
    return iter(results)
