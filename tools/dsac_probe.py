"""Where does one DSAC-variant training round (forward_dsac + backward_dsac) spend its time?  Development aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsac_b200 import engine as E
nb = int(os.environ.get("NB", "16"))
coords, pix, gt_cv, gt_jp = E.synth_frames(max(nb, 1024) if os.environ.get("BIG") else nb)
big = E.Engine(max_frames=1024) if os.environ.get("BIG") else None     # a second, large engine alive (as in tools/sweep_one.py)
eng = E.Engine(max_frames=nb)
for rep in range(int(os.environ.get("REPS", "5"))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fw = eng.forward_dsac(coords[:nb], pix[:nb], gt_jp[:nb], random_draw=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    bw = eng.backward_dsac(nb)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rep %d: forward_dsac %.2f ms  backward_dsac %.2f ms  (%d frames, %.0f refine jobs/frame)" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), nb, bw.n_refine_jobs.mean()), flush=True)
eng.close()
