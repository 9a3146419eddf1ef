"""Hit -> reference-column mapping on the GPU (host wrapper of pa_map_hits).

Drop-in for the per-hit work of the reference's assemble.py main loop with default settings
(/root/reference/lib/assemble.py:1964-1988): ``pretreat_hits`` (:767-873, terminal trimming) and
``read_hit.__init__`` (:13-306 with extend_pos=0).  Consensus assembly, outgroup filtering and
cross-contamination stay on the host as the reference does them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


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


def map_hits(ctx: capi.Context, rows, reads: dict, comps: dict, trim_pos: int = 5, pos_freq: float = 0.2):
    """rows: hit_info.tsv rows (10 fields, any str/int mix); reads: id -> raw nucleotides;
    comps: query name -> composition_table(). Returns a list of dicts (None = dropped hit)."""
    n = len(rows)
    if n == 0:
        return []
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
    cols_cap = int(np.maximum(hmm_to - hmm_from + 1, 0).sum() + n + 8)
    un_cap = int(3 * ali_off[-1] + 3 * n + 8)
    rec = np.zeros(n, dtype=capi.MAPREC_DTYPE)
    base = np.zeros(cols_cap, np.uint8)
    codon = np.zeros(3 * cols_cap, np.uint8)
    unmap = np.zeros(un_cap, np.int32)
    P = capi._ptr
    rc = ctx.L.pa_map_hits(ctx.h, n, model, aligned, P(ali_off), rdcat, P(read_off), P(hmm_from), P(hmm_to), P(ali_from),
                           P(ali_to), P(frame), P(strand), P(comp), P(comp_off), P(ncols_ref), trim_pos, C.c_double(pos_freq),
                           P(rec), P(base), P(codon), P(unmap), cols_cap, un_cap)
    ctx._check(rc, "pa_map_hits")
    out = []
    bb, cb = base.tobytes(), codon.tobytes()
    for h in range(n):
        r = rec[h]
        if not r["keep"]:
            out.append(None)
            continue
        co, nc = int(r["col_off"]), int(r["ncols"])
        out.append({"qstart": int(r["qstart"]), "qend": int(r["qend"]), "ali_from": int(r["ali_from"]), "ali_to": int(r["ali_to"]),
                    "start_trim": int(r["start_trim"]), "end_trim": int(r["end_trim"]),
                    "interval": [int(r["read_from"]), int(r["read_to"]), "+" if r["strand"] > 0 else "-"],
                    "base": bb[co:co + nc].decode(), "codon": cb[3 * co:3 * (co + nc)].decode() if frame[h] >= 0 else None,
                    "unmap": sorted(int(x) for x in unmap[int(r["unmap_off"]):int(r["unmap_off"]) + int(r["n_unmap"])])})
    return out
