"""Multi-GPU path on CPU: world_size-2 gloo run of the frame-sharding host logic (bench.py /
dsac_b200.sharding).  Each rank processes its contiguous shard with frame-keyed sampler streams
(here with the oracle standing in for the GPU engine) -- results must not depend on the split."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from dsac_b200 import engine as E
from dsac_b200.sharding import shard_range, gather_rows
from oracle import oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
n_total = 5
lo, hi = shard_range(n_total, rank, world)
coords, pix, gt_cv, gt_jp = E.synth_frames(hi - lo, frame0=lo)
out = np.zeros((hi - lo, 6))
for i in range(hi - lo):
    cfg = O.default_config(seed=1305 + (lo + i), n_hyps=16, ref_steps=2)
    out[i] = O.forward(cfg, coords[i], pix[i]).ref
full = gather_rows(torch.from_numpy(out), n_total, world)
if rank == 0:
    np.save(sys.argv[2], full.numpy())
dist.barrier()
dist.destroy_process_group()
'''


def test_shard_range_covers_everything():
    from dsac_b200.sharding import shard_range
    for n in (1, 5, 1024, 1000):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_matches_single_process(tmp_path, oracle, engine_mod):
    port = 29500 + (os.getpid() % 500)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    res = tmp_path / "res.npy"
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(res)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = np.load(res)
    coords, pix, _, _ = engine_mod.synth_frames(5)
    for f in range(5):
        cfg = oracle.default_config(seed=1305 + f, n_hyps=16, ref_steps=2)
        assert np.array_equal(got[f], oracle.forward(cfg, coords[f], pix[f]).ref)
