"""GPU probe of the split sampler (sampler_split.cuh) against the monolithic k_sample.

    python tools/k1_split_probe.py [check] [time] [--frames N]

check: same inputs through DSAC_K1_MODE=mono and the default (split) engine; indices / candidate counts must be equal,
       poses equal to rounding.   time: sampler stage alone on N frames x 256 hyp, CUDA events, both modes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsac_b200.engine as E  # noqa: E402


def make(mode, **kw):
    if mode == "mono":
        os.environ["DSAC_K1_MODE"] = "mono"
    else:
        os.environ.pop("DSAC_K1_MODE", None)
    eng = E.Engine(**kw)
    os.environ.pop("DSAC_K1_MODE", None)
    return eng


def check():
    bad = 0
    for T, H, nf in ((1, 256, 3), (8, 256, 3), (3, 64, 3), (64, 64, 3), (1, 256, 64), (1, 64, 5), (5, 256, 7)):
        coords, pix, gt_cv, gt_jp = E.synth_frames(nf, n_streams=T)
        out = {}
        for mode in ("mono", "split"):
            eng = make(mode, max_frames=nf, n_streams=T, n_hyps=H)
            out[mode] = eng.forward(coords, pix, gt_jp)
            if mode == "split":   # a second call exercises the prior taken from the first
                again = eng.forward(coords, pix, gt_jp)
                assert np.array_equal(again.img_idx, out[mode].img_idx) and np.array_equal(again.hyp_pose, out[mode].hyp_pose)
            eng.close()
        a, b = out["mono"], out["split"]
        ok = (np.array_equal(a.img_idx, b.img_idx) and np.array_equal(a.cand_idx, b.cand_idx)
              and np.array_equal(a.n_candidates, b.n_candidates) and np.array_equal(a.status, b.status))
        dp = np.abs(a.hyp_pose - b.hyp_pose).max()
        dr = np.abs(a.ref_pose - b.ref_pose).max()
        print("T=%d H=%d n=%d: indices %s, max |dpose| %.3e, max |dref| %.3e, cands/frame %.0f" % (
            T, H, nf, "EQUAL" if ok else "DIFFER", dp, dr, b.n_candidates.mean()), flush=True)
        if not ok:
            bad += 1
            for f in range(nf):
                if not np.array_equal(a.img_idx[f], b.img_idx[f]):
                    h = int(np.argmax((a.img_idx[f] != b.img_idx[f]).any(1)))
                    print("   frame %d first differing hypothesis %d: mono cand %d idx %s | split cand %d idx %s" % (
                        f, h, a.cand_idx[f][h], a.img_idx[f][h], b.cand_idx[f][h], b.img_idx[f][h]))
                    break
    # capped candidates: sampler exhaustion must be value-encoded the same way
    coords, pix, gt_cv, gt_jp = E.synth_frames(2)
    res = {}
    for mode in ("mono", "split"):
        eng = make(mode, max_frames=2, max_candidates=3000)
        res[mode] = eng.forward(coords, pix, gt_jp)
        eng.close()
    ok = all(np.array_equal(getattr(res["mono"], k), getattr(res["split"], k)) for k in ("img_idx", "cand_idx", "n_candidates", "status"))
    print("max_candidates=3000: %s (status %s)" % ("EQUAL" if ok else "DIFFER", res["split"].status), flush=True)
    bad += 0 if ok else 1
    return bad


def timeit(nf):
    import torch
    coords, pix, gt_cv, gt_jp = E.synth_frames(nf)
    d_coords = torch.from_numpy(coords).cuda(); d_pix = torch.from_numpy(pix).cuda(); d_gt = torch.from_numpy(gt_jp).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for mode in (("split",) if os.environ.get("SPLIT_ONLY") else ("mono", "split")):
        eng = make(mode, max_frames=nf)
        for stages, name in ((E.STAGE_SAMPLE, "sampler"), (E.STAGE_ALL, "step")):
            eng.set_stages(stages)
            eng.set_tail_split(0)
            for _ in range(4):
                eng.forward_device(nf, d_coords.data_ptr(), d_pix.data_ptr(), 0, d_gt.data_ptr(), 0, stream)
            torch.cuda.synchronize()
            tot, reps = 0.0, 10
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                eng.forward_device(nf, d_coords.data_ptr(), d_pix.data_ptr(), 0, d_gt.data_ptr(), 0, stream)
                b.record(); b.synchronize()
                tot += a.elapsed_time(b)
            print("%s %s: %.3f ms / %d frames" % (mode, name, tot / reps, nf), flush=True)
        eng.close()


if __name__ == "__main__":
    nf = 1024
    if "--frames" in sys.argv:
        nf = int(sys.argv[sys.argv.index("--frames") + 1])
    rc = 0
    if "check" in sys.argv or len(sys.argv) == 1:
        rc = check()
    if "time" in sys.argv or len(sys.argv) == 1:
        timeit(nf)
    sys.exit(1 if rc else 0)
