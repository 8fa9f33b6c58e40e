"""GPU check: the event-based candidate-boundary phase of k_sample against the general fixed-point path
(DSAC_K1_A2_GENERIC=1) on the full 1024-frame benchmark batch: sampled cells and candidate numbers must be identical."""
import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(out):
    import torch
    from dsac_b200 import engine as E
    nb = 1024
    coords, pix, gt_cv, gt_jp = E.synth_frames(nb)
    eng = E.Engine(max_frames=nb, write_diffmaps=0)
    r = eng.forward(coords, pix, gt_jp=gt_jp)
    np.savez(out, img_idx=r.img_idx, cand_idx=r.cand_idx, hyp=r.hyp_pose)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1]); sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    a = os.path.join(ROOT, "gpurun_out", "a2_fast.npz"); b = os.path.join(ROOT, "gpurun_out", "a2_generic.npz")
    subprocess.check_call([sys.executable, __file__, a], env=dict(os.environ, DSAC_K1_A2_GENERIC="0"))
    subprocess.check_call([sys.executable, __file__, b], env=dict(os.environ, DSAC_K1_A2_GENERIC="1"))
    A, B = np.load(a), np.load(b)
    ok = all(np.array_equal(A[k], B[k]) for k in ("img_idx", "cand_idx", "hyp"))
    print("a2 fast vs generic identical:", ok, "max cand", int(A["cand_idx"].max()))
    os.remove(a); os.remove(b)
    sys.exit(0 if ok else 1)
