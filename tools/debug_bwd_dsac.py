import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsac_b200 import engine as E
from oracle import oracle as O
H = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nf = 2
coords, pix, gt_cv, gt_jp = E.synth_frames(nf)
eng = E.Engine(max_frames=nf, n_hyps=H)
fw = eng.forward_dsac(coords, pix, gt_jp, random_draw=False)
bw = eng.backward_dsac(nf)
for f in range(nf):
    cfg = O.default_config(seed=1305 + f, n_hyps=H)
    ofw = O.ForwardDsac(cfg)
    ofw.hyp_rvec[:] = fw.hyp_pose[f][:, :3]; ofw.hyp_tvec[:] = fw.hyp_pose[f][:, 3:]
    ofw.img_idx[:] = fw.img_idx[f]; ofw.sf[:] = fw.sf[f]; ofw.ref_pose[:] = fw.ref_pose[f]; ofw.losses[:] = fw.losses[f]
    obw = O.backward_dsac(cfg, coords[f], pix[f], gt_jp[f, :9], gt_jp[f, 9:], ofw)
    d = np.abs(bw.path1[f] - obw.path1)
    print("frame", f, "sel", obw.n_selected, bw.n_selected[f], "jobs", obw.n_refine_jobs, bw.n_refine_jobs[f], "max|p1|", np.abs(obw.path1).max(), "maxdiff", d.max())
    sel = np.where(fw.sf[f] > 1e-4)[0]
    support = {int(c): int(h) for h in sel for c in fw.img_idx[f][h][:3]}
    order = np.argsort(-d.max(1))[:6]
    for c in order:
        print("  cell", c, "diff", d[c], "gpu", bw.path1[f][c], "orc", obw.path1[c], "support of hyp" if c in support else "inlier", support.get(int(c)))
