"""Hit -> reference-column mapping on the GPU (host wrapper of pa_map_hits_ex).

Drop-in for the per-hit work of the reference's assemble.py main loop
(/root/reference/lib/assemble.py:1964-1988): ``pretreat_hits`` (:767-873: short-hit drop, the ``-b``
outgroup-score pre-filter, terminal trimming), ``read_hit.__init__`` (:13-306, incl. ``--extend_pos``) and,
for ``-b -z consensus*``, the per-species accumulation of ``read_hit.consensus`` (:309-633).  Contig
assembly, the outgroup decision on assembled sequences and cross-contamination stay on the host as the
reference does them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

SYMS = "ABCDEFGHIJKLMNOPQRSTUVWXYZ-*"
INT_MAX = 2**31 - 1


def comp_symbol(ch: str) -> int:
    if "A" <= ch <= "Z":
        return ord(ch) - 65
    if "a" <= ch <= "z":
        return ord(ch) - 97
    if ch == "-":
        return 26
    if ch == "*":
        return 27
    return 28


def composition_table(site_composition: dict) -> np.ndarray:
    """get_ref_score()'s site_composition {col: {char: count}} -> int32 [ncols, 32] table."""
    ncols = max(site_composition) if site_composition else 0
    tab = np.zeros((ncols, 32), dtype=np.int32)
    for col, d in site_composition.items():
        for ch, cnt in d.items():
            tab[int(col) - 1, comp_symbol(ch)] += int(cnt)
    return tab


def outgroup_table(outgroup_score: dict, ncols: int) -> np.ndarray:
    """get_ref_score()'s outgroup_score {seqid: {col: score}} (incl. 'PhyloAln_Alt') -> float64 [nseq, ncols],
    NaN where the sequence has no score (column outside its valid ranges)."""
    tab = np.full((len(outgroup_score), ncols), np.nan, dtype=np.float64)
    for s, (_, d) in enumerate(outgroup_score.items()):
        for col, v in d.items():
            tab[s, int(col) - 1] = float(v)
    return tab


def _run(ctx, rows, reads, comps, trim_pos, pos_freq, extend_pos, gencode, out_scores, outgroup_weight, accum):
    n = len(rows)
    genes = sorted({str(r[0]) for r in rows})
    goff, tabs, o = {}, [], 0
    for g in genes:
        goff[g] = (o, comps[g].shape[0])
        tabs.append(comps[g])
        o += comps[g].shape[0]
    comp = np.ascontiguousarray(np.vstack(tabs) if o else np.zeros((1, 32), np.int32), dtype=np.int32)
    model = "".join(str(r[8]) for r in rows).encode()
    aligned = "".join(str(r[9]) for r in rows).encode()
    ali_off = np.zeros(n + 1, np.int64)
    ali_off[1:] = np.cumsum([len(str(r[8])) for r in rows])
    rd = [reads[str(r[1])] for r in rows]
    read_off = np.zeros(n + 1, np.int64)
    read_off[1:] = np.cumsum([len(x) for x in rd])
    rdcat = "".join(rd).encode()
    i32 = lambda k: np.ascontiguousarray([int(r[k]) for r in rows], dtype=np.int32)  # noqa: E731
    hmm_from, hmm_to, ali_from, ali_to = i32(2), i32(3), i32(4), i32(5)
    strand = np.ascontiguousarray([1 if str(r[6]) == "+" else -1 for r in rows], dtype=np.int32)
    frame = np.ascontiguousarray([-1 if str(r[7]) == "None" else int(r[7]) for r in rows], dtype=np.int32)
    comp_off = np.ascontiguousarray([goff[str(r[0])][0] for r in rows], dtype=np.int64)
    ncols_ref = np.ascontiguousarray([goff[str(r[0])][1] for r in rows], dtype=np.int32)
    out_tab = out_off = n_out = None
    if out_scores:
        ooff, parts, o = {}, [], 0
        for g in genes:
            t = np.ascontiguousarray(out_scores[g], dtype=np.float64)
            assert t.shape[1] == goff[g][1]
            ooff[g] = (o, t.shape[0])
            parts.append(t.ravel())
            o += t.size
        out_tab = np.ascontiguousarray(np.concatenate(parts))
        out_off = np.ascontiguousarray([ooff[str(r[0])][0] for r in rows], dtype=np.int64)
        n_out = np.ascontiguousarray([ooff[str(r[0])][1] for r in rows], dtype=np.int32)
    cols_cap = int(np.maximum(hmm_to - hmm_from + 1, 0).sum() + n * (1 + 2 * extend_pos) + 8)
    un_cap = int(3 * ali_off[-1] + 3 * n + 8)
    rec = np.zeros(n, dtype=capi.MAPREC_DTYPE)
    base = np.zeros(cols_cap, np.uint8)
    codon = np.zeros(3 * cols_cap, np.uint8)
    unmap = np.zeros(un_cap, np.int32)
    opts = capi.PaMapOpts(trim_pos, extend_pos, gencode, 1 if out_scores else 0, pos_freq, outgroup_weight)
    P = capi._ptr
    acc_p, keep_alive = None, None
    if accum is not None:
        bins, reps = accum
        keys = []
        for r, b in zip(rows, bins):
            if b is not None and (str(r[0]), b) not in keys:
                keys.append((str(r[0]), b))
        kidx = {k: i for i, k in enumerate(keys)}
        bin_row = np.zeros(len(keys) + 1, np.int64)
        bin_row[1:] = np.cumsum([goff[k[0]][1] for k in keys])
        bin_i = np.ascontiguousarray([-1 if b is None else kidx[(str(r[0]), b)] for r, b in zip(rows, bins)], dtype=np.int32)
        rep_i = np.ascontiguousarray(reps, dtype=np.int32)
        hist = np.zeros((int(bin_row[-1]), 3, 32), np.int32)
        first = np.zeros((int(bin_row[-1]), 3, 32), np.int32)
        qrange = np.empty((max(len(keys), 1), 2), np.int32)
        qrange[:, 0], qrange[:, 1] = INT_MAX, 0
        keep_alive = (keys, bin_row, bin_i, rep_i, hist, first, qrange)
        if keys:
            acc_p = C.pointer(capi.PaAccum(P(bin_i), P(rep_i), len(keys), P(bin_row), P(hist), P(first), P(qrange)))
    rc = ctx.L.pa_map_hits_ex(ctx.h, n, model, aligned, P(ali_off), rdcat, P(read_off), P(hmm_from), P(hmm_to), P(ali_from),
                              P(ali_to), P(frame), P(strand), P(comp), P(comp_off), P(ncols_ref), P(out_tab), P(out_off), P(n_out),
                              C.byref(opts), acc_p, P(rec), P(base), P(codon), P(unmap), cols_cap, un_cap)
    ctx._check(rc, "pa_map_hits_ex")
    return rec, base, codon, unmap, frame, keep_alive


def map_hits(ctx: capi.Context, rows, reads: dict, comps: dict, trim_pos: int = 5, pos_freq: float = 0.2, extend_pos: int = 0,
             gencode: int = 1, out_scores: dict | None = None, outgroup_weight: float = 1.0):
    """rows: hit_info.tsv rows (10 fields, any str/int mix); reads: id -> raw nucleotides;
    comps: query name -> composition_table(); out_scores (the reference's `-b` pre-filter): query name ->
    outgroup_table(). Returns a list of dicts (None = dropped hit)."""
    if len(rows) == 0:
        return []
    rec, base, codon, unmap, frame, _ = _run(ctx, rows, reads, comps, trim_pos, pos_freq, extend_pos, gencode, out_scores,
                                             outgroup_weight, None)
    return _records(rec, base, codon, unmap, frame)


def _records(rec, base, codon, unmap, frame):
    out = []
    bb, cb = base.tobytes(), codon.tobytes()
    for h in range(len(rec)):
        r = rec[h]
        if not r["keep"]:
            out.append(None)
            continue
        co, nc = int(r["col_off"]), int(r["ncols"])
        out.append({"qstart": int(r["qstart"]), "qend": int(r["qend"]), "ali_from": int(r["ali_from"]), "ali_to": int(r["ali_to"]),
                    "start_trim": int(r["start_trim"]), "end_trim": int(r["end_trim"]),
                    "ext_start": int(r["ext_start"]), "ext_end": int(r["ext_end"]),
                    "interval": [int(r["read_from"]), int(r["read_to"]), "+" if r["strand"] > 0 else "-"],
                    "base": bb[co:co + nc].decode(), "codon": cb[3 * co:3 * (co + nc)].decode() if frame[h] >= 0 else None,
                    "unmap": sorted(int(x) for x in unmap[int(r["unmap_off"]):int(r["unmap_off"]) + int(r["n_unmap"])])})
    return out


def consensus_hits(ctx: capi.Context, rows, reads: dict, comps: dict, species, reps, trim_pos: int = 5, pos_freq: float = 0.2,
                   extend_pos: int = 0, gencode: int = 1, out_scores: dict | None = None, outgroup_weight: float = 1.0):
    """The `-b -z consensus*` branch of the reference's hit loop (lib/assemble.py:1973-1974, 1985-1986): every kept
    hit of species sp is folded into ONE read_hit per (alignment, sp) by read_hit.consensus (:309-633).

    species[h]: bin label of hit h (None = skip); reps[h]: number of merged duplicate reads it stands for.
    Returns (per-hit records, {(query, sp): {"qstart", "qend", "seq_comp"}}) where seq_comp has the reference's
    structure and Python dict insertion order: {col: {1: {nt: n}, 2: {...}, 3: {...}}} (codon modes) or
    {col: {nt: n}} (nucleotide mode)."""
    if len(rows) == 0:
        return [], {}
    rec, base, codon, unmap, frame, acc = _run(ctx, rows, reads, comps, trim_pos, pos_freq, extend_pos, gencode, out_scores,
                                               outgroup_weight, (species, reps))
    keys, bin_row, bin_i, _, hist, first, qrange = acc
    out = {}
    for b, key in enumerate(keys):
        hits = np.nonzero((bin_i == b) & (rec["keep"] != 0))[0]
        if len(hits) == 0:
            continue
        nucl = frame[hits[0]] < 0
        r0 = int(bin_row[b])
        H, F = hist[r0:int(bin_row[b + 1])], first[r0:int(bin_row[b + 1])]
        lo, hi = int(qrange[b, 0]), int(qrange[b, 1])
        # insertion order of the columns: by first contributing hit, and inside that hit walk / start extension / end extension
        order = []
        for col in range(lo, hi + 1):
            f = int(F[col - 1].min())
            if f == INT_MAX:
                continue
            r = rec[f]
            q0, q1 = int(r["qstart"]) + int(r["ext_start"]), int(r["qend"]) - int(r["ext_end"])
            within = (0, col) if q0 <= col <= q1 else ((1, -col) if col < q0 else (2, col))
            order.append((f, within, col))
        comp = {}
        for _, _, col in sorted(order):
            def one(k):
                syms = [s for s in range(32) if H[col - 1, k, s]]
                return {SYMS[s]: int(H[col - 1, k, s]) for s in sorted(syms, key=lambda s: int(F[col - 1, k, s]))}
            comp[col] = one(0) if nucl else {1: one(0), 2: one(1), 3: one(2)}
        out[key] = {"qstart": lo, "qend": hi, "seq_comp": comp}
    return _records(rec, base, codon, unmap, frame), out


def ref_score_tables(refseqs: dict, outgroup=(), ingroup=(), sep: str = ".", unknow: str = "N"):
    """The two tables of the mapping kernel straight from the reference alignment: a restatement of get_ref_score()
    (/root/reference/lib/assemble.py:678-764) that skips its per-column dictionaries.

    refseqs: {seqid: aligned sequence} in file order (map/{g}.ref.fas).  Returns (comp, out_names, out_tab):
    comp = composition_table(site_composition), out_tab = outgroup_table(outgroup_score) with out_names its row
    labels (the outgroup sequences found, then 'PhyloAln_Alt').  Float results are produced with the reference's own
    operation order (left-to-right Python sums), so they are bit-identical to the dictionaries' values."""
    ids = list(refseqs)
    if not ids:
        raise ValueError("empty reference alignment")
    ncols = len(refseqs[ids[0]])

    def pick(species):
        found = []
        for sp in species:
            for sid in ids:
                if sid == sp or sid.startswith(sp + sep):
                    found.append(sid)
                    break
        return found
    out_ids = pick(outgroup) if outgroup else list(ids)
    if ingroup:
        in_ids = pick(ingroup)
    elif len(out_ids) == len(ids):
        in_ids = list(ids)
    else:
        in_ids = [sid for sid in ids if sid not in out_ids]
    # valid ranges: runs of residues whose internal gaps are <= 5 long and that hold > 5 non-gap... exactly :706-723
    valid = {}
    for sid in ids:
        s = refseqs[sid]
        ok = np.zeros(ncols, dtype=bool)
        gap = nogap = 0
        i = -1
        for i, ch in enumerate(s):
            if ch == "-" or ch == unknow:
                gap += 1
                if gap > 5:
                    if nogap > 5:
                        ok[i - gap - nogap + 1:i - gap + 1] = True
                    nogap = 0
            else:
                if nogap > 0 and 0 < gap <= 5:
                    nogap += gap
                nogap += 1
                gap = 0
        if nogap > 5:
            ok[max(i - gap - nogap + 1, 0):i - gap + 1] = True
        valid[sid] = ok
    comp = np.zeros((ncols, 32), dtype=np.int32)
    for sid in in_ids:
        s, ok = refseqs[sid], valid[sid]
        for i in np.nonzero(ok)[0]:
            comp[i, comp_symbol(s[i])] += 1
    # dictionary keys are the literal characters: two characters that share a table slot ('?' and '#', say) would be
    # separate keys in the reference; alignments only hold letters, '-', '*' and the unknown symbol
    totals = comp.sum(axis=1)
    out_tab = np.full((len(out_ids) + 1, ncols), np.nan, dtype=np.float64)
    means = []
    for r, sid in enumerate(out_ids):
        s, ok = refseqs[sid], valid[sid]
        fr = []
        for i in np.nonzero(ok)[0]:
            cnt = int(comp[i, comp_symbol(s[i])])
            out_tab[r, i] = float(cnt)
            fr.append(cnt / int(totals[i]))      # ZeroDivisionError as in the reference when no ingroup residue is valid here
        means.append(sum(fr) / len(fr))
    aver = min(means)
    for i in range(ncols):
        out_tab[len(out_ids), i] = int(totals[i]) * aver
    return comp, out_ids + ["PhyloAln_Alt"], out_tab


# ---------------------------------------------------------------------------------------------
# merge_hits(): folding hits into contigs (lib/assemble.py:889-954; callers :995, :1074, :1157, :1198, :1214, :1580, :1766)
# ---------------------------------------------------------------------------------------------
def _absorb_intervals(have: list, new: list, codon: bool):
    """--ref_split_len bookkeeping of merge_hits (:897-914): a fragment that continues an interval of the same strand (and,
    with codons, the same reading frame) extends it, anything else is appended."""
    for s, e, strand in new:
        for iv in have:
            if iv[2] != strand or (codon and (s - iv[0]) % 3 != 0):
                continue
            if strand == "+":
                if iv[0] <= s <= iv[1] + 1 and e >= iv[1]:
                    iv[1] = e
                    break
            elif e <= iv[1] and e + 1 >= iv[0] and s <= iv[0]:
                iv[0] = s
                break
        else:
            have.append([s, e, strand])


def merge_hits_chains(ctx: capi.Context, chains, mol_type: str = "dna", split_len=None, fill_gap=False):
    """Left fold of merge_hits() over every chain: chains[j] = [item, item, ...] in merge order, an item being a dict
    {"qstart", "qend", "seq_comp", "seqids": {id: [[start, end, strand], ...]}, "count": {id: 1} | {"too much": n}} with
    seq_comp in the reference's structure ({col: {nt: n}} for mol_type 'dna', {col: {1: {..}, 2: {..}, 3: {..}}} otherwise).

    The per-column histograms -- the bulk of the data -- are summed on the GPU (pa_merge_hists: segmented fold that also
    returns which item brought each symbol first, i.e. the insertion order of the reference's dictionaries); qend, the
    target-sequence intervals, the counts and the fill_gap columns are host bookkeeping.  Returns one merged item per chain,
    equal to the reference's result including dict order."""
    codon = mol_type != "dna"
    K = 3 if codon else 1
    fill_sym = comp_symbol(fill_gap) if fill_gap else None
    # ---- flatten: items (plus one synthetic gap item after every hit that opens a gap) -> dense [column][3][32] blocks
    chain_off, q0s, ncs, blocks, metas = [0], [], [], [], []
    for chain in chains:
        if not chain:
            raise ValueError("empty chain")
        flat, end = [], None
        for n, it in enumerate(chain):
            flat.append(it["seq_comp"])
            if n and fill_gap and it["qstart"] - end > 1:
                cell = {fill_gap: 1}
                flat.append({c: ({1: dict(cell), 2: dict(cell), 3: dict(cell)} if codon else dict(cell)) for c in range(end + 1, it["qstart"])})
            end = it["qend"] if n == 0 else max(end, it["qend"])
        metas.append(flat)
        for comp in flat:
            cols = list(comp)
            lo, hi = (min(cols), max(cols)) if cols else (1, 0)
            q0s.append(lo)
            ncs.append(hi - lo + 1)
            blocks.append(comp)
        chain_off.append(len(q0s))
    item_row = np.zeros(len(q0s) + 1, np.int64)
    item_row[1:] = np.cumsum(ncs)
    hist_in = np.zeros((max(int(item_row[-1]), 1), 3, 32), np.int32)
    key_in = np.zeros_like(hist_in)
    for i, comp in enumerate(blocks):
        for c, d in comp.items():
            row = int(item_row[i]) + c - q0s[i]
            for k in range(K):
                for rank, (ch, cnt) in enumerate((d[k + 1] if codon else d).items()):
                    s = comp_symbol(ch)
                    if s >= 28:
                        raise ValueError(f"unexpected symbol {ch!r} in seq_comp")
                    hist_in[row, k, s] += cnt
                    key_in[row, k, s] = rank
    chain_nc = np.array([max((q0s[i] + ncs[i] - 1 for i in range(chain_off[j], chain_off[j + 1])), default=0) for j in range(len(chains))],
                        dtype=np.int32)
    chain_nc = np.maximum(chain_nc, 0)
    chain_row = np.zeros(len(chains) + 1, np.int64)
    chain_row[1:] = np.cumsum(chain_nc)
    hist = np.zeros((max(int(chain_row[-1]), 1), 3, 32), np.int32)
    who, key = np.zeros_like(hist), np.zeros_like(hist)
    P = capi._ptr
    coff_a, q0_a, nc_a = (np.ascontiguousarray(x, dtype=np.int32) for x in (chain_off, q0s, ncs))   # keep alive across the call
    rc = ctx.L.pa_merge_hists(ctx.h, len(chains), P(coff_a), P(q0_a), P(nc_a), P(item_row), P(hist_in), P(key_in), P(chain_row), P(chain_nc),
                              P(hist), P(who), P(key))
    ctx._check(rc, "pa_merge_hists")
    # ---- rebuild the merged items
    out = []
    for j, chain in enumerate(chains):
        first = chain[0]
        new = {"qstart": first["qstart"], "qend": max(it["qend"] for it in chain)}
        # column order: the first item's columns, then every later item's new columns in its own order (gap columns after the hit)
        order = {}
        for comp in metas[j]:
            for c in comp:
                order.setdefault(c, None)
        H, W, Kk = (x[int(chain_row[j]):int(chain_row[j + 1])] for x in (hist, who, key))
        seq_comp = {}
        for c in order:
            def one(k):
                syms = np.flatnonzero(W[c - 1, k] != INT_MAX)
                return {SYMS[s]: int(H[c - 1, k, s]) for s in sorted(syms, key=lambda s: (int(W[c - 1, k, s]), int(Kk[c - 1, k, s])))}
            seq_comp[c] = {1: one(0), 2: one(1), 3: one(2)} if codon else one(0)
        new["seq_comp"] = seq_comp
        # target-sequence bookkeeping (:893-935, :941-949)
        seqids = {k: [list(iv) for iv in v] for k, v in first["seqids"].items()}
        count = dict(first["count"])
        end = first["qend"]
        for it in chain[1:]:
            last_id = None
            if "too much" not in count:
                for sid, ivs in it["seqids"].items():
                    last_id = sid
                    have = seqids.setdefault(sid, [])
                    if split_len:
                        _absorb_intervals(have, ivs, codon)
                    else:
                        have.extend(list(iv) for iv in ivs)
                    count[sid] = 1
            else:
                count["too much"] += it["count"]["too much"]
            if fill_gap and it["qstart"] - end > 1 and "too much" not in count:
                ivs = seqids[last_id]     # the reference looks at the last id of the merged-in hit (:942)
                if ivs[-1][2] == "+":
                    if ivs[-1][0] - ivs[-2][1] <= 1:
                        ivs[-2][1] = ivs[-1][1]
                        ivs.pop()
                elif ivs[-2][0] - ivs[-1][1] <= 1:
                    ivs[-2][0] = ivs[-1][0]
                    ivs.pop()
            end = max(end, it["qend"])
        new["seqids"], new["count"] = seqids, count
        out.append(new)
    return out
