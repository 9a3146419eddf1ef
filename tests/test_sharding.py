"""N > 1 host path on CPU: two ranks (launched with torchrun, gloo rendezvous -- the TEST's plumbing; the product's
gather uses files, no torch and no NCCL) shard the reads by chunk, search their shard (the oracle stands in for the GPU
on this CPU-only box), gather hit records on rank 0 and apply the global report (Z, domZ, order).  Must equal the
single-process result -- no collective on the data path."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from oracle import oracle as O
from phyloaln_b200 import capi, search, synth
from tests import helpers
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)
hmms = [helpers.calibrated_hmm(f"s{i}", M, seed=300 + i, bg_mix=0.1) for i, M in enumerate([50, 90])]
reads = synth.random_reads(240, 120, rng)
synth.plant_reads(reads, hmms, 0.3, rng, sub_rate=0.02)
rl = [r.tobytes().decode() for r in reads]
ids = [f"read{i}_PhAlTag_SP_fastx1" for i in range(len(rl))]
a, b = search.shard_bounds(len(rl), world, rank)
recs, model, aligned = [], [], []
frames = [p for r in rl[a:b] for _, p in O.six_frames(r, 1)]
for hi, h in enumerate(hmms):
    for d in O.Profile.from_hmm(h).search(frames, nthreads=2, want_trace=False)[0]:
        r = np.zeros(1, dtype=capi.HIT_DTYPE)[0]
        r["hmm"] = hi
        for f in ("ndom_in_seq", "dom_idx", "ienv", "jenv", "iali", "jali", "hmmfrom", "hmmto", "dom_bits", "dom_lnP", "seq_bits", "seq_lnP"):
            r[f] = d[f]
        r["target"] = d["target"] + a * 6          # local -> global target index
        recs.append(r); model.append(d["model"]); aligned.append(d["aligned"])
hits = np.array(recs, dtype=capi.HIT_DTYPE) if recs else np.zeros(0, dtype=capi.HIT_DTYPE)
hits, model, aligned = search.gather_hits_files(hits, model, aligned, rank, world, sys.argv[2] + '.gather', dst=0)
dist.barrier()
if rank == 0:
    tname = lambda t: ids[t // 6] + search.frame_suffix(t % 6, 6, True)[0]
    rep = search.report(hits, model, aligned, tname, len(rl) * 6, search.SearchOptions(), len(hmms))
    rows = {h: search.hit_rows(f"g{h}", rep[h], ids, 6, True) for h in rep}
    json.dump(rows, open(sys.argv[2], "w"))
dist.destroy_process_group()
'''


def _run(world, out, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29511 + world), str(script), ROOT, str(out)]
    subprocess.run(cmd, check=True, env=env, timeout=600, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return json.load(open(out))


def test_two_rank_gloo_shards_equal_single_process(tmp_path):
    one = _run(1, tmp_path / "one.json", tmp_path)
    two = _run(2, tmp_path / "two.json", tmp_path)
    assert one == two
    assert sum(len(v) for v in one.values()) > 20
