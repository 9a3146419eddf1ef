"""TEST-ONLY stand-in for Bio.AlignIO.read(path, "stockholm"), written from Biopython's documented behaviour
(StockholmIterator): exactly one alignment per file, blocks concatenated per identifier in order of first appearance,
'.' of a sequence row read as '-', `#=GC RF` exposed as _per_col_annotations['reference_annotation'], rows and column
annotations of unequal length rejected.  Used by tests/golden/make_refcols_golden.py to run the reference's own
reference-column code (lib/aln.py:750-792); never imported by the product."""

_GC = {"RF": "reference_annotation", "SS_cons": "secondary_structure", "SA_cons": "surface_accessibility", "TM_cons": "transmembrane",
       "PP_cons": "posterior_probability", "LI_cons": "ligand_binding", "AS_cons": "active_site", "MM": "model_mask"}


class _Seq(str):
    pass


class _Record:
    def __init__(self, sid, seq):
        self.id = sid
        self.name = sid.split("/")[0]
        self.seq = _Seq(seq)


class _Alignment(list):
    _per_col_annotations = None

    def get_alignment_length(self):
        return len(self[0].seq) if self else 0


def read(handle, fmt):
    if fmt != "stockholm":
        raise ValueError(f"the stand-in reads Stockholm only, not {fmt}")
    own = isinstance(handle, str)
    fh = open(handle) if own else handle
    try:
        lines = fh.read().split("\n")
    finally:
        if own:
            fh.close()
    if not lines or lines[0].strip() != "# STOCKHOLM 1.0":
        raise ValueError("Did not find STOCKHOLM header")
    order, seqs, gc = [], {}, {}
    i, done = 1, False
    while i < len(lines):
        line = lines[i].strip()
        i += 1
        if line == "# STOCKHOLM 1.0":
            raise ValueError("Found second STOCKHOLM header before end of the first alignment")
        if line == "//":
            done = True
            break
        if not line:
            continue
        if line[0] != "#":
            parts = [x.strip() for x in line.split(" ", 1)]
            if len(parts) != 2:
                raise ValueError("Could not split line into identifier and sequence:\n" + line)
            if parts[0] not in seqs:
                order.append(parts[0])
                seqs[parts[0]] = ""
            seqs[parts[0]] += parts[1].replace(".", "-")
        elif line[:5] == "#=GC ":
            feature, text = line[5:].strip().split(None, 1)
            gc[feature] = gc.get(feature, "") + text.strip()
    if done and any(x.strip() for x in lines[i:]):
        raise ValueError("More than one record found in handle")
    if not order:
        raise ValueError("No records found in handle")
    n = len(seqs[order[0]])
    aln = _Alignment()
    for sid in order:
        if len(seqs[sid]) != n:
            raise ValueError("Sequences have different lengths, or repeated identifier")
        aln.append(_Record(sid, seqs[sid]))
    ann = {}
    for k, v in gc.items():
        if len(v) != n:
            raise ValueError(f"{k} length {len(v)}, expected {n}")
        ann[_GC.get(k, "GC:" + k)] = v
    aln._per_col_annotations = ann
    return aln
