"""Freeze outputs of the REFERENCE'S OWN code (oracle/_ref: core/cnn_softam.h, train_ransac_softam.cpp, ... compiled
unmodified against oracle/shim) as golden fixtures, so that the pinning travels to boxes without /root/reference.

    python tests/golden/make_ref_golden.py        ->  tests/golden/ref_golden.npz

Inputs are the synthetic frames of SURVEY.md section 8(d) (dsac_synth_frames, data seed 20170721, sampler seed 1305);
every case stores the frame's inputs too, so the consumers do not depend on the generator staying unchanged.
Consumers: tests/test_oracle_vs_ref.py (oracle vs fixtures, CPU) and tests/test_gpu_ref_golden.py (CUDA engine vs fixtures).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dsac_b200.engine as E  # noqa: E402
from oracle import ref as R  # noqa: E402

# (name, H, T, frames, with_gradient)
CASES = [("c1_h64_t1", 64, 1, (0, 1, 2), True), ("c2_h256_t1", 256, 1, (0, 5), True), ("c3_h256_t8", 256, 8, (3,), True),
         ("h64_t3", 64, 3, (7,), False)]


def main():
    out = {}
    for name, H, T, frames, grad in CASES:
        for g in frames:
            coords, pix, gt_cv, gt_jp = E.synth_frames(1, frame0=g, n_streams=T)
            key = "%s_f%d" % (name, g)
            with tempfile.TemporaryDirectory() as d:
                R.write_dataset(d, "training", gt_jp)
                Rg, tg = R.read_pose(d, os.path.join("training", "synth", "poses", "frame-000000.pose.txt"))
                cfg = R.config(n_hyps=H, n_threads=T, frame=g)
                fw = R.forward(cfg, coords[0], Rg, tg)
                assert np.array_equal(fw.pix, pix[0]) and np.array_equal(fw.est_obj, coords[0])
                rec = dict(H=H, T=T, frame=g, coords=coords[0], pix=pix[0], gt_R=Rg, gt_t=tg, img_idx=fw.img_idx,
                           hyp_rvec=fw.hyp_rvec, hyp_tvec=fw.hyp_tvec, scores=fw.scores, sf=fw.sf, entropy=fw.entropy, avg=fw.avg,
                           ref=fw.ref, inlier_map=fw.inlier_map, n_perm_steps=fw.n_perm_steps, loss=fw.loss, rot_err=fw.rot_err,
                           t_err=fw.t_err, correct=fw.correct,
                           diffmap_rows=fw.diffmaps[:: max(1, H // 8)])   # every (H/8)-th row of the HxN matrix
                if grad:
                    cfg1 = R.config(n_hyps=H, n_threads=T, frame=0)   # the one-frame dataset's frame 0 carries global index g
                    cfg1.frame = g
                    dl, loss, sog = R.train_round(cfg1, d, coords[0], args=("-rI", str(H)))
                    assert abs(loss - fw.loss) < 1e-12
                    fa = R.factors(cfg, want_dref_dobj=False)
                    rec.update(dloss_dobj=dl, score_out_grads=sog, dloss_dref=fa["dloss_dref"], dref_dhyp=fa["dref_dhyp"])
            for k, v in rec.items():
                out[key + "/" + k] = np.asarray(v)
            print(key, "loss %.6f" % fw.loss, "correct", fw.correct, flush=True)
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
