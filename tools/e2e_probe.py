"""e2e timing probe: dsac_forward with pinned host buffers for different pipeline chunk counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsac_b200 import engine as E
nb = 1024
coords, pix, gt_cv, gt_jp = E.synth_frames(nb)
def pinned(a):
    t = torch.from_numpy(a).pin_memory(); return t, t.numpy()
k1, hc = pinned(coords); k2, hp = pinned(pix); k3, hg = pinned(gt_jp)
# raw copy bandwidth
d = torch.empty(hc.nbytes + hp.nbytes, dtype=torch.uint8, device="cuda")
src = torch.empty(hc.nbytes + hp.nbytes, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(src, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("H2D %.1f MB in %.3f ms -> %.1f GB/s" % (src.numel() / 1e6, dt * 1e3, src.numel() / dt / 1e9))
for chunks in (1, 2, 3, 4):
    os.environ["DSAC_PIPE_CHUNKS"] = str(chunks)
    eng = E.Engine(max_frames=nb)
    out = E.ForwardResult(nb, 256, False)
    keep = []
    for name in ("ref_pose", "avg_pose", "sf", "scores", "entropy", "loss", "rot_err", "t_err", "correct", "status", "n_candidates"):
        tt, arr = pinned(getattr(out, name)); keep.append(tt); setattr(out, name, arr); setattr(out.raw, name, arr.ctypes.data)
    for name in ("hyp_pose", "img_idx", "cand_idx", "diffmaps", "inlier_map", "ref_steps_done", "n_perm_steps"):
        setattr(out.raw, name, None)
    for _ in range(3): eng.forward(hc, hp, hg, out=out)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): eng.forward(hc, hp, hg, out=out)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 100
    print("chunks", chunks, "e2e %.3f ms/step -> %.2f Mhyp/s" % (ms, nb * 256 / ms / 1e3), "acc", out.correct.mean())
    eng.close()
