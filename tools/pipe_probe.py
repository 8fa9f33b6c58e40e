"""Development aid: the chained generator (k1_pipe, DSAC_K1_PIPE=1) against the ordinary path: identical results? time per step?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsac_b200 import engine as E
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for nf in [int(x) for x in (sys.argv[1:] or ["16", "64", "128", "140"])]:
    coords, pix, gt_cv, gt_jp = E.synth_frames(nf)
    d_c, d_p, d_g = torch.from_numpy(coords).cuda(), torch.from_numpy(pix).cuda(), torch.from_numpy(gt_jp).cuda()
    st = torch.cuda.current_stream().cuda_stream
    res, ms = {}, {}
    for pipe in ("1", "0"):
        os.environ["DSAC_K1_PIPE"] = pipe
        eng = E.Engine(max_frames=nf)
        for _ in range(3): eng.forward_device(nf, d_c.data_ptr(), d_p.data_ptr(), 0, d_g.data_ptr(), 0, st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.forward_device(nf, d_c.data_ptr(), d_p.data_ptr(), 0, d_g.data_ptr(), 0, st); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        ms[pipe] = float(np.mean(ts))
        res[pipe] = eng.fetch(nf)
        eng.close()
    same = all(np.array_equal(getattr(res["1"], k), getattr(res["0"], k)) for k in ("img_idx", "cand_idx", "n_candidates", "hyp_pose", "scores", "ref_pose", "inlier_map", "status"))
    print("frames %4d: pipe %.3f ms, ordinary %.3f ms per step; identical results: %s; status sum %d" % (nf, ms["1"], ms["0"], same, int(res["1"].status.sum())), flush=True)
