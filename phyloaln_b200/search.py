"""Host side of the B200 search step: option parsing, sharding, report thresholds and the
PhyloAln-facing outputs.

Mirrors, for the default ``-j hmmer`` mode, what the reference does around the ``hmmsearch``
child process (/root/reference/lib/aln.py):

* ``assemble()`` :650-674 builds ``hmmsearch -o .. --tblout .. --cpu N <params> g.hmm all.target.fasta``;
  here :func:`parse_hmmsearch_parameters` understands the same pass-through parameter list and
  :class:`B200Searcher` runs all profiles against a resident read chunk on the GPU.
* HMMER applies the E-value thresholds with ``Z`` = number of target sequences and ``domZ`` =
  number of reported sequences after all targets are seen and sorts the report
  (A.10); :func:`report` does that on the gathered hit records (the only global step, host side,
  no collective on the data path).
* ``extract_reads()`` :482-570 turns the text report into ``map/{g}.hit_info.tsv`` rows and
  ``map/{g}.targets.fas``; :func:`hit_rows` / :func:`write_step_outputs` produce the same files
  and ``ok/`` markers directly from the hit records.
"""
from __future__ import annotations

import math
import os
import threading
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .hmmio import HMM, read_hmm


class SearchParameterError(ValueError):
    pass


@dataclass
class SearchOptions:
    """hmmsearch options that affect the hit set (reference: args.hmmsearch_parameters, lib/aln.py:878)."""
    E: float = 10.0
    domE: float = 10.0
    T: float | None = None
    domT: float | None = None
    incE: float = 0.01
    incdomE: float = 0.01
    Z: float | None = None
    domZ: float | None = None
    F1: float = 0.02
    F2: float = 1e-3
    F3: float = 1e-5
    nobias: bool = False
    nonull2: bool = False
    max: bool = False
    ignored: list = field(default_factory=list)


_FLOAT_OPTS = {"-E": "E", "--domE": "domE", "-T": "T", "--domT": "domT", "--incE": "incE", "--incdomE": "incdomE",
               "-Z": "Z", "--domZ": "domZ", "--F1": "F1", "--F2": "F2", "--F3": "F3"}
_FLAG_OPTS = {"--nobias": "nobias", "--nonull2": "nonull2", "--max": "max"}
_IGNORED_WITH_ARG = {"--cpu", "--seed", "-o", "-A", "--tblout", "--domtblout", "--pfamtblout", "--textw", "--tformat"}
_IGNORED_FLAGS = {"--acc", "--noali", "--notextw", "--cut_ga", "--cut_nc", "--cut_tc"}


def parse_hmmsearch_parameters(params) -> SearchOptions:
    """Parse PhyloAln's pass-through list (e.g. [' -E', '1e-5']; leading blanks as in lib/main.py:11-12)."""
    o = SearchOptions()
    toks = []
    for p in params or []:
        toks.extend(str(p).split())
    i = 0
    while i < len(toks):
        t = toks[i]
        if t in _FLOAT_OPTS:
            if i + 1 >= len(toks):
                raise SearchParameterError(f"hmmsearch option {t} needs a value")
            try:
                setattr(o, _FLOAT_OPTS[t], float(toks[i + 1]))
            except ValueError:
                raise SearchParameterError(f"bad value for {t}: {toks[i + 1]}") from None
            i += 2
        elif t in _FLAG_OPTS:
            setattr(o, _FLAG_OPTS[t], True)
            i += 1
        elif t in _IGNORED_WITH_ARG:
            o.ignored.append(t)
            i += 2
        elif t in _IGNORED_FLAGS:
            if t.startswith("--cut"):
                raise SearchParameterError(f"{t}: bit-score cutoffs (GA/TC/NC) are not supported by the B200 search path")
            o.ignored.append(t)
            i += 1
        else:
            raise SearchParameterError(f"unsupported hmmsearch option for the B200 search path: {t}")
    return o


def frame_suffix(frame: int, nframes: int, translated: bool):
    """Target-name suffix and (strand, start_pos) of frame index f (order of lib/aln.py:241-258)."""
    if translated:
        nfwd = 1 if nframes == 1 else 3
        if nframes == 1 or frame < nfwd:
            return f"_pos{frame + 1}", "+", frame
        return f"_rev_pos{frame - nfwd + 1}", "rev", frame - nfwd
    return ("", "+", None) if frame == 0 else ("_rev", "rev", None)


def report(hits: np.ndarray, model, aligned, target_names, Z: float, opts: SearchOptions, n_hmm: int, pp=None):
    """Apply reporting thresholds and HMMER's report order per profile.

    hits: structured array (capi.HIT_DTYPE) gathered from all shards, ``target`` already global.
    target_names: callable target_index -> full target name (with _pos/_rev suffix).
    Returns {hmm_index: [domain dict, ...]} in report order.
    """
    out = {h: [] for h in range(n_hmm)}
    if len(hits) == 0:
        return out
    order = np.lexsort((hits["dom_idx"], hits["target"], hits["hmm"]))
    hs = hits[order]
    key = hs["hmm"].astype(np.int64) * (1 << 40) + hs["target"].astype(np.int64)
    starts = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    ends = np.r_[starts[1:], len(hs)]
    Zeff = float(opts.Z) if opts.Z is not None else float(Z)
    per_hmm: dict[int, list] = {}
    for s, e in zip(starts, ends):
        h = int(hs["hmm"][s])
        lnP = float(hs["seq_lnP"][s])
        score = float(hs["seq_bits"][s])
        if opts.T is not None:
            ok = score >= opts.T
        else:
            ok = math.exp(lnP) * Zeff <= opts.E
        if ok:
            per_hmm.setdefault(h, []).append((s, e))
    for h, seqs in per_hmm.items():
        domZ = float(opts.domZ) if opts.domZ is not None else float(len(seqs))
        names = {s: target_names(int(hs["target"][s])) for s, _ in seqs}
        # sort key -lnP descending (E-value ascending); ties by name (strcmp)
        seqs.sort(key=lambda se: (float(hs["seq_lnP"][se[0]]), names[se[0]].encode()))
        rows = []
        for s, e in seqs:
            for g in range(s, e):
                if opts.domT is not None:
                    ok = float(hs["dom_bits"][g]) >= opts.domT
                else:
                    ok = math.exp(float(hs["dom_lnP"][g])) * domZ <= opts.domE
                if not ok:
                    continue
                og = int(order[g])
                rows.append({"hmm": h, "target": int(hs["target"][g]), "name": names[s], "dom_idx": int(hs["dom_idx"][g]),
                             "hmmfrom": int(hs["hmmfrom"][g]), "hmmto": int(hs["hmmto"][g]), "alifrom": int(hs["iali"][g]),
                             "alito": int(hs["jali"][g]), "envfrom": int(hs["ienv"][g]), "envto": int(hs["jenv"][g]),
                             "seq_bits": float(hs["seq_bits"][g]), "seq_evalue": math.exp(float(hs["seq_lnP"][g])) * Zeff,
                             "seq_bias": float(hs["seqbias"][g]) / math.log(2.0),
                             "dom_bits": float(hs["dom_bits"][g]), "dom_cevalue": math.exp(float(hs["dom_lnP"][g])) * domZ,
                             "dom_ievalue": math.exp(float(hs["dom_lnP"][g])) * Zeff,
                             "dom_bias": float(hs["dombias"][g]) / math.log(2.0), "ndom": int(hs["ndom_in_seq"][g]),
                             "oasc": float(hs["oasc"][g]), "domZ": domZ, "Z": Zeff,
                             "model": model[og], "aligned": aligned[og], "pp": pp[og] if pp is not None else None})
        out[h] = rows
    return out


def hit_rows(query: str, rows, ids, nframes: int, translated: bool):
    """hit_info.tsv rows exactly as extract_reads() writes them (lib/aln.py:544-558)."""
    out = []
    for r in rows:
        read, frame = divmod(r["target"], nframes)
        _, strand, start_pos = frame_suffix(frame, nframes, translated)
        out.append((query, ids[read], r["hmmfrom"], r["hmmto"], r["alifrom"], r["alito"], strand, start_pos,
                    r["model"].upper().replace(".", "-"), r["aligned"].upper().replace(".", "-")))
    return out


def write_step_outputs(outdir: str, query: str, rows, ids, raw_seqs, nframes: int, translated: bool, log_lines=()):
    """map/{g}.hit_info.tsv, map/{g}.targets.fas, map/{g}.hmmsearch.log and both ok/ markers
    (lib/aln.py:500-501, 558-561, 569, 663, 673). raw_seqs: callable read_index -> raw nucleotides."""
    os.makedirs(os.path.join(outdir, "map"), exist_ok=True)
    os.makedirs(os.path.join(outdir, "ok"), exist_ok=True)
    seen = set()
    with open(os.path.join(outdir, "map", query + ".hit_info.tsv"), "w") as tsv, \
            open(os.path.join(outdir, "map", query + ".targets.fas"), "w") as fas:
        for r, row in zip(rows, hit_rows(query, rows, ids, nframes, translated)):
            tsv.write("\t".join(str(x) for x in row) + "\n")
            read = r["target"] // nframes
            if row[1] not in seen:
                fas.write(f">{row[1]}\n{raw_seqs(read)}\n")
                seen.add(row[1])
    with open(os.path.join(outdir, "map", query + ".hmmsearch.log"), "w") as log:
        for line in log_lines:
            log.write(line + "\n")
    open(os.path.join(outdir, "ok", f"hmmsearch_{query}.ok"), "w").close()
    open(os.path.join(outdir, "ok", f"extract_reads_{query}.ok"), "w").close()


def write_tblout(path: str, query: str, rows):
    """Minimal HMMER --tblout table for diffing against a real hmmsearch."""
    with open(path, "w") as fh:
        fh.write("#                                                               --- full sequence ---- --- best 1 domain ---- --- domain number estimation ----\n")
        fh.write("# target name        accession  query name           accession    E-value  score  bias   E-value  score  bias   exp reg clu  ov env dom rep inc description of target\n")
        fh.write("#------------------- ---------- -------------------- ---------- --------- ------ ----- --------- ------ -----   --- --- --- --- --- --- --- --- ---------------------\n")
        seen = {}
        for r in rows:
            seen.setdefault(r["target"], []).append(r)
        for doms in seen.values():
            best = max(doms, key=lambda d: d["dom_bits"])
            r = doms[0]
            fh.write(f"{r['name']:<20s} -          {query:<20s} -          {r['seq_evalue']:9.2g} {r['seq_bits']:6.1f} {r['seq_bias']:5.1f} "
                     f"{best['dom_ievalue']:9.2g} {best['dom_bits']:6.1f} {best['dom_bias']:5.1f}   {float(r['ndom']):3.1f} {r['ndom']:3d}   0   0 "
                     f"{r['ndom']:3d} {r['ndom']:3d} {len(doms):3d} {len(doms):3d} -\n")


def write_domtblout(path: str, query: str, qlen: int, rows, target_lengths=None):
    """HMMER --domtblout table: one line per reported domain, 22 columns + description."""
    with open(path, "w") as fh:
        fh.write("#                                                                            --- full sequence --- -------------- this domain -------------   hmm coord   ali coord   env coord\n")
        fh.write("# target name        accession   tlen query name           accession   qlen   E-value  score  bias   #  of  c-Evalue  i-Evalue  score  bias  from    to  from    to  from    to  acc description of target\n")
        fh.write("#------------------- ---------- ----- -------------------- ---------- ----- --------- ------ ----- --- --- --------- --------- ------ ----- ----- ----- ----- ----- ----- ----- ---- ---------------------\n")
        seqs = []
        for r in rows:
            if seqs and seqs[-1][0]["target"] == r["target"]:
                seqs[-1].append(r)
            else:
                seqs.append([r])
        for doms in seqs:
            for n, d in enumerate(doms, 1):
                tlen = target_lengths(d["target"]) if target_lengths else 0
                acc = d.get("oasc", 0.0) / (1.0 + abs(d["envto"] - d["envfrom"]))
                fh.write(f"{d['name']:<20s} -          {tlen:5d} {query:<20s} -          {qlen:5d} {d['seq_evalue']:9.2g} {d['seq_bits']:6.1f} "
                         f"{d['seq_bias']:5.1f} {n:3d} {len(doms):3d} {d['dom_cevalue']:9.2g} {d['dom_ievalue']:9.2g} {d['dom_bits']:6.1f} "
                         f"{d['dom_bias']:5.1f} {d['hmmfrom']:5d} {d['hmmto']:5d} {d['alifrom']:5d} {d['alito']:5d} {d['envfrom']:5d} "
                         f"{d['envto']:5d} {acc:4.2f} -\n")


def shard_bounds(nreads: int, world: int, rank: int):
    """Contiguous read range of one rank (reads shard by chunk; profiles are replicated)."""
    per = (nreads + world - 1) // world
    return min(nreads, rank * per), min(nreads, (rank + 1) * per)


def gather_hits_distributed(hits: np.ndarray, model, aligned, dst: int = 0):
    """Host-side gather of per-rank hit lists (the only cross-rank step; no collective on the DP path).

    Uses torch.distributed object gather so it works with the nccl and gloo backends alike.
    Returns (hits, model, aligned) on rank dst, (None, None, None) elsewhere."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = (hits.tobytes(), list(model), list(aligned))
    bucket = [None] * world if rank == dst else None
    dist.gather_object(payload, bucket, dst=dst)
    if rank != dst:
        return None, None, None
    parts = [np.frombuffer(b[0], dtype=capi.HIT_DTYPE) for b in bucket]
    allhits = np.concatenate(parts) if parts else np.zeros(0, dtype=capi.HIT_DTYPE)
    return allhits, [m for b in bucket for m in b[1]], [m for b in bucket for m in b[2]]


class B200Searcher:
    """All profiles x one read set on one or more B200s (one context per GPU, reads sharded by chunk)."""

    def __init__(self, hmms, devices=(0,), options: SearchOptions | None = None):
        self.opts = options or SearchOptions()
        self.hmms = [h if isinstance(h, HMM) else read_hmm(h) for h in hmms]
        self.ctxs = []
        for d in devices:
            c = capi.Context(d)
            c.set_opts(F1=self.opts.F1, F2=self.opts.F2, F3=self.opts.F3, do_bias=not self.opts.nobias,
                       do_null2=not self.opts.nonull2, do_max=self.opts.max)
            for h in self.hmms:
                c.add_hmm(h)
            c.commit()
            self.ctxs.append(c)

    def close(self):
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def search_reads(self, nt: np.ndarray, off: np.ndarray, mol_type: str = "prot", gencode: int = 1, no_reverse: bool = False,
                     codon: bool = False, chunk_reads: int = 1 << 18, want_pp: bool = False):
        """Search concatenated ASCII reads. Returns (hits, model, aligned, Z, nframes, translated); with want_pp the PP
        annotation lines of the alignments are left in ``self.last_pp`` (same order as hits).

        ``mol_type`` follows PhyloAln's -m: 'dna*' searches nucleotide targets, everything else translates
        (lib/aln.py:241-258). Reads are cut into chunks; chunk c goes to GPU c mod n (contexts run in
        parallel threads; ctypes releases the GIL). Hit target indices are global."""
        translated = not mol_type.startswith("dna")
        nreads = len(off) - 1
        bounds = list(range(0, nreads, chunk_reads)) + [nreads]
        results = [None] * (len(bounds) - 1)
        nframes_box = [1]
        errors = []

        def search_range(ctx, a, b, out):
            """Search reads [a, b); on a capacity error (-3) split the range (the library has no fallback)."""
            sub_off = (off[a:b + 1] - off[a]).astype(np.uint64)
            sub_nt = nt[int(off[a]):int(off[b])]
            if len(sub_nt) == 0:
                sub_nt = np.zeros(1, dtype=np.uint8)
            ctx.load_reads(sub_nt, sub_off, moltype=0 if translated else 1, gencode=gencode, no_reverse=no_reverse,
                           codon_only=codon)
            nf = ctx.num_frames()
            nframes_box[0] = nf
            try:
                hits, model, aligned = ctx.search()
            except capi.PaError as e:
                if "(-3)" in str(e) and b - a > 1:
                    mid = (a + b) // 2
                    search_range(ctx, a, mid, out)
                    search_range(ctx, mid, b, out)
                    return
                raise
            pp = ctx.hit_pp(hits) if want_pp else [None] * len(hits)
            hits = hits.copy()
            hits["target"] += a * nf
            out.append((hits, model, aligned, pp))

        def work(ci, ctx):
            try:
                for c in range(ci, len(bounds) - 1, len(self.ctxs)):
                    parts = []
                    search_range(ctx, bounds[c], bounds[c + 1], parts)
                    results[c] = (np.concatenate([p[0] for p in parts]), [m for p in parts for m in p[1]],
                                  [m for p in parts for m in p[2]], [m for p in parts for m in p[3]])
            except Exception as e:  # surfaced after join
                errors.append(e)

        threads = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate(self.ctxs)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        nframes = nframes_box[0]
        res = [r for r in results if r is not None]
        hits = np.concatenate([r[0] for r in res]) if res else np.zeros(0, dtype=capi.HIT_DTYPE)
        model = [m for r in res for m in r[1]]
        aligned = [m for r in res for m in r[2]]
        self.last_pp = [m for r in res for m in r[3]] if want_pp else None
        Z = nreads * nframes
        return hits, model, aligned, Z, nframes, translated

    def search_and_report(self, nt, off, ids, **kw):
        hits, model, aligned, Z, nframes, translated = self.search_reads(nt, off, **kw)
        self.last_Z = Z

        def tname(t):
            read, frame = divmod(t, nframes)
            return ids[read] + frame_suffix(frame, nframes, translated)[0]
        rep = report(hits, model, aligned, tname, Z, self.opts, len(self.hmms), pp=self.last_pp)
        return rep, nframes, translated
