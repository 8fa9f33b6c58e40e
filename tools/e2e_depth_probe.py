"""How many engines in flight does the host-buffer loop (dsac_forward_submit / dsac_forward_wait) need?  Wall-clock ms per
1024-frame step for 1..4 engines, pinned host buffers, results read back every step."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dsac_b200 import engine as E

nf, H, steps = int(os.environ.get("NB", "1024")), 256, int(os.environ.get("STEPS", "20"))
coords, pix, gt_cv, gt_jp = E.synth_frames(nf)
keep = []
def pinned(a):
    t = torch.from_numpy(a).pin_memory(); keep.append(t); return t.numpy()
h_coords, h_pix, h_gt = pinned(coords), pinned(pix), pinned(gt_jp)
def host_result():
    out = E.ForwardResult(nf, H, False)
    for name in ("ref_pose", "avg_pose", "sf", "scores", "entropy", "loss", "rot_err", "t_err", "correct", "status", "n_candidates"):
        arr = pinned(getattr(out, name)); setattr(out, name, arr); setattr(out.raw, name, arr.ctypes.data)
    for name in ("hyp_pose", "img_idx", "cand_idx", "diffmaps", "inlier_map", "ref_steps_done", "n_perm_steps"):
        setattr(out.raw, name, None)
    return out
for depth in (1, 2, 3, 4):
    engs = [E.Engine(max_frames=nf) for _ in range(depth)]
    outs = [host_result() for _ in range(depth)]
    def run(n):
        for i in range(n):
            e = engs[i % depth]; e.forward_wait(); e.forward_submit(h_coords, h_pix, h_gt, out=outs[i % depth])
        for e in engs: e.forward_wait()
    run(3 * depth + 2); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(steps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("engines in flight %d: %.3f ms / step, %.1f M hyp/s" % (depth, 1e3 * dt / steps, nf * H * steps / dt / 1e6), flush=True)
    for e in engs: e.close()
