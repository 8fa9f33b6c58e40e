"""Development aid: the speculative first round (k1_spec) against the one-CTA generator on many single frames (every frame is
another seed): identical results, and how often the speculation is abandoned.   python tools/spec_stress.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsac_b200 import engine as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
os.environ["DSAC_K1_SPEC"] = "1"; a = E.Engine(max_frames=1)
os.environ["DSAC_K1_SPEC"] = "0"; b = E.Engine(max_frames=1)
stitched, abandoned, diff = 0, 0, 0
for f in range(n):
    coords, pix, gt_cv, gt_jp = E.synth_frames(1, frame0=f)
    ra = a.forward(coords, pix, gt_jp, frame0=f)
    w = a.spec_result()[0]
    rb = b.forward(coords, pix, gt_jp, frame0=f)
    stitched += w > 0; abandoned += w == 0
    same = all(np.array_equal(getattr(ra, k), getattr(rb, k)) for k in ("img_idx", "cand_idx", "n_candidates", "hyp_pose", "scores", "ref_pose", "status"))
    diff += not same
print("frames %d: speculation stitched %d, abandoned %d; results differing from the one-CTA generator: %d" % (n, stitched, abandoned, diff))
a.close(); b.close()
