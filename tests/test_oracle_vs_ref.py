"""Pins the oracle (oracle/dsac_oracle.cpp, the CPU restatement every GPU parity test is checked against) to the
REFERENCE'S OWN CODE.

oracle/_ref is /root/reference/core/{cnn_softam.h, maxloss.h, Hypothesis.cpp, types.h, thread_rand.cpp, properties.cpp,
read_data.cpp, dataset.h, lua_calls.h, test_ransac_softam.cpp, train_ransac_softam.cpp} compiled UNMODIFIED against API
shims (oracle/shim: OpenCV containers with the calib3d numerics forwarded to the oracle's restated primitives, a Lua
C-API stand-in whose "score CNN" is the closed-form soft-inlier score, png++) -- see oracle/Makefile.  What is pinned
is therefore everything above the OpenCV boundary: the sampling loop and its RNG use, reprojection-error maps, softmax /
soft-argmax, the refinement loop and its stop rules, evaluation, every finite-difference factor of the backward pass
and the gradient assembly inside the reference's own main(), including the quirks Q1, Q2, Q4, Q6, Q7, Q8, Q10
(SURVEY.md section 7).  The OpenCV boundary itself is pinned to cv2 4.13 in tests/test_oracle_golden.py.

Two layers: (1) live, where /root/reference exists (this container); (2) against tests/golden/ref_golden.npz, frozen
from oracle/_ref by tests/golden/make_ref_golden.py, everywhere (the GPU box has no reference tree).

Tolerances: integer work (sampled indices, inlier maps, permutations, step counts) bit-exact; poses, scores, softmax,
losses and all gradient factors 1e-12 relative (measured: 0 .. 3e-14).
"""
import os
import tempfile

import numpy as np
import pytest

from oracle import ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")
live = pytest.mark.skipif(not R.available(), reason="the reference tree (/root/reference) is not present on this box")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def check_forward(o, r, H):
    assert np.array_equal(o.img_idx, r["img_idx"])
    assert rel(o.hyp_rvec, r["hyp_rvec"]) <= 1e-12 and rel(o.hyp_tvec, r["hyp_tvec"]) <= 1e-12
    assert rel(o.scores, r["scores"]) <= 1e-12 and np.abs(o.sf - r["sf"]).max() <= 1e-12
    assert abs(o.entropy - float(r["entropy"])) <= 1e-12
    assert rel(o.avg, r["avg"]) <= 1e-12 and rel(o.ref, r["ref"]) <= 1e-12
    assert np.array_equal(o.inlier_map, r["inlier_map"]) and o.n_perm_steps == int(r["n_perm_steps"])
    assert abs(o.loss - float(r["loss"])) <= 1e-10 and abs(o.rot_err - float(r["rot_err"])) <= 1e-10
    assert abs(o.t_err - float(r["t_err"])) <= 1e-9 and o.correct == int(r["correct"])


# ------------------------------------------------------------------------------------------------ live
@live
def test_rng_contract_is_the_references(oracle, engine_mod):
    """stochasticSubSample through the reference's ThreadRand (thread_rand.cpp:59-81, cnn_softam.h:283-309)."""
    for seed in (1305, 1306, 99):
        ref_pix = R.stochastic_subsample(seed)
        assert np.array_equal(ref_pix, oracle.stochastic_subsample(seed))
        assert np.array_equal(ref_pix, engine_mod.stochastic_subsample(seed))


@live
@pytest.mark.parametrize("T,H,frames", [(1, 64, (0, 1, 2)), (1, 256, (4,)), (8, 256, (2,)), (3, 64, (6,)), (8, 64, (9, 10))])
def test_forward_equals_reference_process_image(oracle, engine_mod, T, H, frames):
    """processImage (cnn_softam.h:960-1180) with OMP_NUM_THREADS = T  <->  orc_forward with n_streams = T (configs 1, 2)."""
    O, E = oracle, engine_mod
    for g in frames:
        coords, pix, gt_cv, gt_jp = E.synth_frames(1, frame0=g, n_streams=T)
        r = R.forward(R.config(n_hyps=H, n_threads=T, frame=g), coords[0], gt_jp[0, :9], gt_jp[0, 9:])
        assert np.array_equal(r.pix, pix[0]) and np.array_equal(r.est_obj, coords[0])
        o = O.forward(O.default_config(seed=1305 + g * T, n_hyps=H, n_streams=T), coords[0], pix[0], gt_jp[0, :9], gt_jp[0, 9:])
        assert o.n_fragile == 0
        ref = dict(img_idx=r.img_idx, hyp_rvec=r.hyp_rvec, hyp_tvec=r.hyp_tvec, scores=r.scores, sf=r.sf, entropy=r.entropy, avg=r.avg,
                   ref=r.ref, inlier_map=r.inlier_map, n_perm_steps=r.n_perm_steps, loss=r.loss, rot_err=r.rot_err, t_err=r.t_err,
                   correct=r.correct)
        check_forward(o, ref, H)
        assert np.abs(o.diffmaps - r.diffmaps).max() == 0.0          # getDiffMap, cnn_softam.h:319-362
        assert np.array_equal(o.pixel_idxs, r.pixel_idxs)             # std::shuffle permutations, quirk Q2


@live
@pytest.mark.parametrize("T,H,g", [(1, 64, 1), (8, 256, 3)])
def test_backward_factors_equal_the_references(oracle, engine_mod, T, H, g):
    """dLossMax, dRefineObj (x skip, Q8), dRefineHyp (unit mix, Q7), dPNP by the reference's own functions (config 3)."""
    O, E = oracle, engine_mod
    coords, pix, gt_cv, gt_jp = E.synth_frames(1, frame0=g, n_streams=T)
    cfg = R.config(n_hyps=H, n_threads=T, frame=g)
    R.forward(cfg, coords[0], gt_jp[0, :9], gt_jp[0, 9:])
    oc = O.default_config(seed=1305 + g * T, n_hyps=H, n_streams=T)
    o = O.forward(oc, coords[0], pix[0], gt_jp[0, :9], gt_jp[0, 9:])
    ob = O.backward(oc, coords[0], pix[0], gt_jp[0, :9], gt_jp[0, 9:], o)
    fa = R.factors(cfg)
    assert rel(ob.dloss_dref, fa["dloss_dref"]) <= 1e-12
    assert rel(ob.dref_dhyp, fa["dref_dhyp"]) <= 1e-12
    assert rel(ob.dref_dobj, fa["dref_dobj"]) <= 1e-12 and np.abs(fa["dref_dobj"]).max() > 0
    assert rel(ob.dpnp, fa["dpnp"]) <= 1e-12


@live
@pytest.mark.parametrize("T,H,g", [(1, 64, 0), (1, 256, 2), (8, 256, 5)])
def test_training_round_gradient_equals_the_references_main(oracle, engine_mod, T, H, g):
    """One round of main() of train_ransac_softam.cpp (forward, paths I and II, the x-major dScore columns of quirk Q4, the
    assembly of lines 288-394) on a one-frame dataset on disk; dLoss_dObj is captured where the driver hands it to the
    coordinate CNN (:412).  The ground truth goes through the reference's 7-Scenes pose reader (read_data.cpp:69-133)."""
    O, E = oracle, engine_mod
    coords, pix, gt_cv, gt_jp = E.synth_frames(1, frame0=g, n_streams=T)
    with tempfile.TemporaryDirectory() as d:
        R.write_dataset(d, "training", gt_jp)
        Rg, tg = R.read_pose(d, os.path.join("training", "synth", "poses", "frame-000000.pose.txt"))
        assert np.abs(Rg.reshape(-1) - gt_jp[0, :9]).max() < 1e-6 and np.abs(tg - gt_jp[0, 9:]).max() < 1e-3   # float round trip
        dl, loss, sog = R.train_round(R.config(n_hyps=H, n_threads=T, frame=g), d, coords[0], args=("-rI", str(H)))
    oc = O.default_config(seed=1305 + g * T, n_hyps=H, n_streams=T)
    o = O.forward(oc, coords[0], pix[0], Rg, tg)
    ob = O.backward(oc, coords[0], pix[0], Rg, tg, o)
    assert abs(o.loss - loss) <= 1e-10
    assert rel(ob.score_grads, sog) <= 1e-10
    assert rel(ob.dloss_dobj, dl) <= 1e-12 and np.abs(dl).max() > 0


@live
def test_test_driver_logs_equal_the_references_main(oracle, engine_mod):
    """main() of test_ransac_softam.cpp over a 12-frame trajectory (config 5 in small): its per-frame log (loss, entropy,
    errors, pose converted back to the 7-Scenes convention, :161-223) and its summary line (:251-263) against the same
    quantities from the oracle, compared at the 6 significant digits the reference prints."""
    O, E = oracle, engine_mod
    n, H = 12, 64
    coords, pix, gt_cv, gt_jp = E.synth_frames(n, traj=True)
    tr = np.array([0.25, -0.5, 1.0])
    with tempfile.TemporaryDirectory() as d:
        R.write_dataset(d, "test", gt_jp, translation=tr)
        gts = [R.read_pose(d, os.path.join("test", "synth", "poses", "frame-%06d.pose.txt" % i)) for i in range(n)]
        logs = R.run_test_main(R.config(n_hyps=H), d, coords, args=("-rI", str(H)))
    per_frame = [ln.split() for ln in logs["ransac_test_errors_obj_model_init.net_rdraw1_softam.txt"].strip().splitlines()]
    summary = logs["ransac_test_loss_obj_model_init.net_rdraw1_softam.txt"].split()
    assert len(per_frame) == n
    losses, ents, rots, ts, corr = [], [], [], [], []
    for i in range(n):
        o = O.forward(O.default_config(seed=1305 + i, n_hyps=H), coords[i], pix[i], gts[i][0], gts[i][1])
        Rj, tj = O.cv2our(o.ref[:3], o.ref[3:])
        M = np.eye(4); M[:3, :3] = Rj; M[:3, 3] = tj
        P = np.linalg.inv(M) @ np.diag([1.0, -1.0, -1.0, 1.0])
        want = [o.loss, o.entropy, o.t_err, o.rot_err] + list(O.rodrigues_inv(P[:3, :3])) + list(P[:3, 3] / 1000.0 + tr)
        got = [float(v) for v in per_frame[i]]
        assert len(got) == 10
        for w, gv in zip(want, got):
            assert abs(w - gv) <= 1e-5 * max(1.0, abs(w)), (i, want, got)
        losses.append(o.loss); ents.append(o.entropy); rots.append(o.rot_err); ts.append(o.t_err); corr.append(o.correct)
    want = [np.mean(corr), np.mean(losses), np.std(losses), np.mean(ents), np.std(ents), sorted(rots)[n // 2], sorted(ts)[n // 2]]
    for w, gv in zip(want, [float(v) for v in summary]):
        assert abs(w - gv) <= 1e-5 * max(1.0, abs(w))


# ------------------------------------------------------------------------------------------------ frozen fixtures
def _cases():
    z = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def test_oracle_matches_the_reference_fixtures(oracle):
    """The same comparison against outputs of oracle/_ref frozen in tests/golden/ref_golden.npz (runs on any box)."""
    O = oracle
    z, names = _cases()
    assert len(names) >= 7
    for name in names:
        g = {k.split("/")[1]: z[k] for k in z.files if k.startswith(name + "/")}
        H, T, f = int(g["H"]), int(g["T"]), int(g["frame"])
        oc = O.default_config(seed=1305 + f * T, n_hyps=H, n_streams=T)
        o = O.forward(oc, g["coords"], g["pix"], g["gt_R"], g["gt_t"])
        check_forward(o, g, H)
        assert np.abs(o.diffmaps[:: max(1, H // 8)] - g["diffmap_rows"]).max() == 0.0
        if "dloss_dobj" in g:
            ob = O.backward(oc, g["coords"], g["pix"], g["gt_R"], g["gt_t"], o)
            assert rel(ob.dloss_dobj, g["dloss_dobj"]) <= 1e-12
            assert rel(ob.score_grads, g["score_out_grads"]) <= 1e-10
            assert rel(ob.dloss_dref, g["dloss_dref"]) <= 1e-12 and rel(ob.dref_dhyp, g["dref_dhyp"]) <= 1e-12


@live
def test_pose_file_reader_equals_the_references(engine_mod, tmp_path):
    """the host mirror's 7-Scenes pose reader (dsac_b200/host/read_data.cpp) against the
    reference's own readData + Hypothesis(info) (core/read_data.cpp:69-133) on pose files with a scene offset."""
    import subprocess
    E = engine_mod
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "apps"), "-s", "host_selftest"])
    coords, pix, gt_cv, gt_jp = E.synth_frames(5, traj=True)
    d = str(tmp_path)
    R.write_dataset(d, "test", gt_jp, translation=[0.5, -1.25, 2.0])
    for i in range(5):
        rel_path = os.path.join("test", "synth", "poses", "frame-%06d.pose.txt" % i)
        Rg, tg = R.read_pose(d, rel_path)
        out = subprocess.run([os.path.join(root, "apps", "host_selftest"), "pose", rel_path], cwd=d, capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stdout + out.stderr
        v = np.array([float(x) for x in out.stdout.split()])
        assert np.abs(v[:9] - Rg.reshape(-1)).max() <= 2e-7 and np.abs(v[9:] - tg).max() <= 1e-3   # float arithmetic on both sides
        assert np.abs(v[:9] - gt_jp[i, :9]).max() <= 1e-6 and np.abs(v[9:] - gt_jp[i, 9:]).max() <= 5e-3


# ------------------------------------------------------------------------------------------------ DSAC / RANSAC variant
@live
@pytest.mark.parametrize("T,H,g,random_draw", [(1, 64, 0, True), (1, 64, 3, False), (4, 32, 2, True)])
def test_dsac_variant_forward_equals_reference_process_image(oracle, engine_mod, T, H, g, random_draw):
    """processImage of core/cnn.h:1028-1257 (draw :102-126, refinement of every hypothesis :1155-1228, expectedMaxLoss
    :137-151) by the reference's own code  <->  orc_forward_dsac."""
    O, E = oracle, engine_mod
    coords, pix, gt_cv, gt_jp = E.synth_frames(1, frame0=g, n_streams=T)
    r = R.forward_dsac(R.config(n_hyps=H, n_threads=T, frame=g), coords[0], gt_jp[0, :9], gt_jp[0, 9:], random_draw)
    o = O.forward_dsac(O.default_config(seed=1305 + g * T, n_hyps=H, n_streams=T), coords[0], pix[0], gt_jp[0, :9], gt_jp[0, 9:], random_draw)
    assert np.array_equal(o.img_idx, r.img_idx) and o.hyp_idx == r.hyp_idx and o.correct == r.correct
    assert np.array_equal(o.inlier_maps, r.inlier_maps)
    assert rel(o.hyp_rvec, r.hyp_rvec) <= 1e-12 and rel(o.hyp_tvec, r.hyp_tvec) <= 1e-12
    assert np.abs(o.sf - r.sf).max() <= 1e-12 and rel(o.ref_pose, r.ref_pose) <= 1e-12
    assert np.abs(o.losses - r.losses).max() <= 1e-9 and abs(o.expected_loss - r.expected_loss) <= 1e-10
    assert abs(o.rot_err - r.rot_err) <= 1e-9 and abs(o.t_err - r.t_err) <= 1e-9


@live
@pytest.mark.parametrize("T,H,g", [(1, 32, 1), (4, 48, 6)])
def test_dsac_variant_gradient_equals_the_references_main(oracle, engine_mod, T, H, g):
    """One round of main() of core/train_ransac.cpp (expectation of the loss: sum_h sf_h dLossMax dRefine_h + dSMScore,
    lines 304-381) on a one-frame dataset  <->  orc_backward_dsac."""
    O, E = oracle, engine_mod
    coords, pix, gt_cv, gt_jp = E.synth_frames(1, frame0=g, n_streams=T)
    with tempfile.TemporaryDirectory() as d:
        R.write_dataset(d, "training", gt_jp)
        Rg, tg = R.read_pose(d, os.path.join("training", "synth", "poses", "frame-000000.pose.txt"))
        dl, loss = R.train_round_dsac(R.config(n_hyps=H, n_threads=T, frame=g), d, coords[0], args=("-rI", str(H)))
    oc = O.default_config(seed=1305 + g * T, n_hyps=H, n_streams=T)
    o = O.forward_dsac(oc, coords[0], pix[0], Rg, tg, True)
    ob = O.backward_dsac(oc, coords[0], pix[0], Rg, tg, o)
    assert abs(o.expected_loss - loss) <= 1e-10
    assert rel(ob.dloss_dobj, dl) <= 1e-10 and np.abs(dl).max() > 0
