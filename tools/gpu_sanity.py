"""Quick GPU sanity + timing script (development aid; the graded checks live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dsac_b200 import engine as E
from oracle import oracle as O

nf = 3
coords, pix, gt_cv, gt_jp = E.synth_frames(nf)
for T in (1, 8):
    eng = E.Engine(max_frames=nf, n_streams=T)
    res = eng.forward(coords, pix, gt_jp, want_diffmaps=True)
    for f in range(nf):
        cfg = O.default_config(seed=1305 + f * T, n_streams=T)
        fw = O.forward(cfg, coords[f], pix[f], gt_jp[f, :9], gt_jp[f, 9:])
        idx_ok = (fw.img_idx == res.img_idx[f]).all()
        cand_ok = (fw.cand_idx == res.cand_idx[f]).all()
        pose_d = np.abs(np.concatenate([fw.hyp_rvec, fw.hyp_tvec / 1000], 1) - res.hyp_pose[f] / np.array([1, 1, 1, 1000, 1000, 1000])).max()
        dm_d = np.abs(fw.diffmaps - res.diffmaps[f]).max()
        sc_d = np.abs(fw.scores - res.scores[f]).max() / np.abs(fw.scores).max()
        sf_d = np.abs(fw.sf - res.sf[f]).max()
        avg_d = np.abs(fw.avg - res.avg_pose[f]).max()
        ref_d = np.abs(fw.ref - res.ref_pose[f]).max()
        print(f"T={T} f={f} idx_ok={idx_ok} cand_ok={cand_ok} ncand {fw.n_candidates}/{res.n_candidates[f]} pose {pose_d:.2e} dm {dm_d:.2e} "
              f"score {sc_d:.2e} sf {sf_d:.2e} avg {avg_d:.2e} ref {ref_d:.2e} steps {fw.ref_steps_done}/{res.ref_steps_done[f]} "
              f"imap_ok={(fw.inlier_map == res.inlier_map[f]).all()} loss {fw.loss:.4f}/{res.loss[f]:.4f} rot {fw.rot_err:.4f}/{res.rot_err[f]:.4f} status {res.status[f]}")
    eng.close()

# timing on a batch
nb = int(os.environ.get("NB", "1024"))
coords, pix, gt_cv, gt_jp = E.synth_frames(nb)
eng = E.Engine(max_frames=nb)
dc = torch.from_numpy(coords).cuda(); dp = torch.from_numpy(pix).cuda(); dg = torch.from_numpy(gt_jp).cuda()
st = torch.cuda.current_stream().cuda_stream
for name, mask in (("sample", 1), ("score", 2), ("refine", 4 | 8), ("all", 15)):
    eng.set_stages(15); eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st); torch.cuda.synchronize()
    eng.set_stages(mask)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    ev0.record()
    for _ in range(reps):
        eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    print(f"{name}: {ms:.3f} ms per {nb} frames -> {nb*256/ms*1e3/1e6:.2f} Mhyp/s")
res = eng.fetch(nb)
print("mean candidates/frame", res.n_candidates.mean(), "correct", res.correct.mean(), "median rot", np.median(res.rot_err), "median t", np.median(res.t_err), "status", np.bincount(res.status))
# CPU baseline
cfg = O.default_config()
ns = 16
t = O.bench_forward(cfg, coords[:ns], pix[:ns], n_threads=os.cpu_count(), with_refine=True)
print(f"oracle: {ns} frames in {t:.2f}s on {os.cpu_count()} threads -> {ns*256/t/1e3:.2f} khyp/s")
