"""Where a dsac_forward call (host pinned buffers, 1024 frames) spends its time: DSAC_TRACE=1 prints H2D / kernels / D2H."""
import os, sys, time
os.environ["DSAC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsac_b200 import engine as E
nb = 1024
coords, pix, gt_cv, gt_jp = E.synth_frames(nb)
def pinned(a):
    t = torch.from_numpy(a).pin_memory(); return t, t.numpy()
k1, hc = pinned(coords); k2, hp = pinned(pix); k3, hg = pinned(gt_jp)
eng = E.Engine(max_frames=nb)
out = E.ForwardResult(nb, 256, False)
keep = []
for name in ("ref_pose", "avg_pose", "sf", "scores", "entropy", "loss", "rot_err", "t_err", "correct", "status", "n_candidates"):
    tt, arr = pinned(getattr(out, name)); keep.append(tt); setattr(out, name, arr); setattr(out.raw, name, arr.ctypes.data)
for name in ("hyp_pose", "img_idx", "cand_idx", "diffmaps", "inlier_map", "ref_steps_done", "n_perm_steps"):
    setattr(out.raw, name, None)
for _ in range(6):
    t0 = time.perf_counter(); eng.forward(hc, hp, hg, out=out); print("python call %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
