"""One library build (DSAC_B200_LIB) measured on the GPU: stage-isolated kernel times, the whole step with the tail
split off / on, the end-to-end loops, a DSAC-variant training round, and checksums of the results.  Prints one JSON
line.  Development aid for tools/sweep.py -- not part of the product or of bench.py."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import dsac_b200.engine as E  # noqa: E402

H, NF = 256, int(os.environ.get("SWEEP_FRAMES", "1024"))
STEPS = int(os.environ.get("SWEEP_STEPS", "10"))


def main():
    lib = E.load()
    has_split = hasattr(lib, "dsac_set_tail_split")
    coords, pix, gt_cv, gt_jp = E.synth_frames(NF)
    eng = E.Engine(max_frames=NF)
    stream = torch.cuda.current_stream().cuda_stream
    d_c, d_p, d_g = torch.from_numpy(coords).cuda(), torch.from_numpy(pix).cuda(), torch.from_numpy(gt_jp).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def split(e, m):
        if has_split:
            e.set_tail_split(m)

    def step():
        eng.forward_device(NF, d_c.data_ptr(), d_p.data_ptr(), 0, d_g.data_ptr(), 0, stream)

    def timed(fn, k):
        tot = []
        for _ in range(k):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            tot.append(a.elapsed_time(b))
        return float(np.mean(tot)), float(np.min(tot))

    out = {"lib": os.environ.get("DSAC_B200_LIB", "default"), "has_split": has_split, "frames": NF}
    eng.set_stages(E.STAGE_ALL)
    split(eng, 0)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    out["step_split0_ms"] = timed(step, STEPS)
    res0 = eng.fetch(NF)
    if has_split:
        split(eng, 1)
        step(); step()
        out["step_split1_ms"] = timed(step, STEPS)
        res1 = eng.fetch(NF)
        out["split_identical"] = bool(all(np.array_equal(getattr(res0, k), getattr(res1, k)) for k in ("img_idx", "ref_pose", "scores", "inlier_map", "loss")))
    split(eng, 0)
    for name, mask in (("k_sample", E.STAGE_SAMPLE), ("k_score", E.STAGE_SCORE), ("k_refine", E.STAGE_REFINE | E.STAGE_EVAL)):
        eng.set_stages(E.STAGE_ALL); step()
        eng.set_stages(mask); step(); torch.cuda.synchronize()
        out[name + "_ms"] = timed(step, STEPS)
    eng.set_stages(E.STAGE_ALL)
    out["checksum"] = {"ref_pose": float(np.abs(res0.ref_pose).sum()), "loss": float(res0.loss.sum()), "inlier_map": int(res0.inlier_map.sum()),
                       "img_idx": int(res0.img_idx.sum()), "scores": float(res0.scores.sum()), "correct": float(res0.correct.mean()),
                       "steps_done": int(res0.ref_steps_done.sum()) if res0.ref_steps_done is not None else None}

    # ---- end to end with host buffers
    def pinned(a):
        t_ = torch.from_numpy(a).pin_memory()
        return t_, t_.numpy()
    keep = []
    k1, h_c = pinned(coords); k2, h_p = pinned(pix); k3, h_g = pinned(gt_jp)

    def host_result():
        o = E.ForwardResult(NF, H, False)
        for name in ("ref_pose", "avg_pose", "sf", "scores", "entropy", "loss", "rot_err", "t_err", "correct", "status", "n_candidates"):
            tt, arr = pinned(getattr(o, name)); keep.append(tt)
            setattr(o, name, arr); setattr(o.raw, name, arr.ctypes.data)
        for name in ("hyp_pose", "img_idx", "cand_idx", "diffmaps", "inlier_map", "ref_steps_done", "n_perm_steps"):
            setattr(o.raw, name, None)
        return o
    o0, o1 = host_result(), host_result()
    eng_b = E.Engine(max_frames=NF)
    engs, outs = (eng, eng_b), (o0, o1)

    def run_sync(k):
        for _ in range(k):
            eng.forward(h_c, h_p, h_g, frame0=0, out=o0)

    def run_pipe(k):
        for i in range(k):
            engs[i & 1].forward_wait()
            engs[i & 1].forward_submit(h_c, h_p, h_g, frame0=0, out=outs[i & 1])
        engs[0].forward_wait(); engs[1].forward_wait()

    for mode in ((0, 1) if has_split else (0,)):
        split(eng, mode)
        run_sync(3); torch.cuda.synchronize()
        t0 = time.perf_counter(); run_sync(STEPS); torch.cuda.synchronize()
        out["e2e_sync_split%d_ms" % mode] = (time.perf_counter() - t0) * 1e3 / STEPS
    for mode in ((1, 2) if has_split else (1,)):
        split(eng, mode); split(eng_b, mode)
        run_pipe(4); torch.cuda.synchronize()
        t0 = time.perf_counter(); run_pipe(2 * STEPS); torch.cuda.synchronize()
        out["e2e_pipe_submitsplit%d_ms" % (mode == 2)] = (time.perf_counter() - t0) * 1e3 / (2 * STEPS)
    # three engines in flight
    eng_c = E.Engine(max_frames=NF)
    o2 = host_result()
    engs3, outs3 = (eng, eng_b, eng_c), (o0, o1, o2)
    for e_ in engs3:
        split(e_, 1)

    def run_pipe3(k):
        for i in range(k):
            engs3[i % 3].forward_wait()
            engs3[i % 3].forward_submit(h_c, h_p, h_g, frame0=0, out=outs3[i % 3])
        for e_ in engs3:
            e_.forward_wait()
    run_pipe3(6); torch.cuda.synchronize()
    t0 = time.perf_counter(); run_pipe3(3 * STEPS); torch.cuda.synchronize()
    out["e2e_pipe3_ms"] = (time.perf_counter() - t0) * 1e3 / (3 * STEPS)
    eng_c.close()
    eng_b.close()

    # ---- DSAC-variant round (refinement of all hypotheses + its backward: k_refine throughput)
    nb = 16
    eng5 = E.Engine(max_frames=nb)
    eng5.forward_dsac(coords[:nb], pix[:nb], gt_jp[:nb], random_draw=True)
    bw = eng5.backward_dsac(nb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fw = eng5.forward_dsac(coords[:nb], pix[:nb], gt_jp[:nb], random_draw=True)
        bw = eng5.backward_dsac(nb)
    torch.cuda.synchronize()
    out["dsac_round_ms"] = (time.perf_counter() - t0) * 1e3 / (3 * nb)
    out["dsac_checksum"] = {"expected_loss": float(np.sum(fw.expected_loss)), "grad_abs": float(np.abs(bw.dloss_dobj).sum()), "refine_jobs": float(bw.n_refine_jobs.mean())}
    eng5.close()
    # ---- softam training round
    eng3 = E.Engine(max_frames=64)
    eng3.forward(coords[:64], pix[:64], gt_jp[:64]); g = eng3.backward(coords[:64], pix[:64], gt_jp[:64], full=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng3.forward(coords[:64], pix[:64], gt_jp[:64]); g = eng3.backward(coords[:64], pix[:64], gt_jp[:64], full=False)
    torch.cuda.synchronize()
    out["softam_round_ms"] = (time.perf_counter() - t0) * 1e3 / (3 * 64)
    out["softam_grad_abs"] = float(np.abs(g.dloss_dobj).sum())
    eng3.close()
    eng.close()
    print("SWEEP " + json.dumps(out))


if __name__ == "__main__":
    main()
