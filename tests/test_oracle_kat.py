"""Known-answer / analytic-identity tests of the oracle (SURVEY.md section 4 consequence (ii))."""
import numpy as np
import pytest


def _frame(engine_mod, f=0, **kw):
    coords, pix, gt_cv, gt_jp = engine_mod.synth_frames(1, frame0=f, **kw)
    return coords[0], pix[0], gt_cv[0], gt_jp[0]


def test_p3p_recovers_generating_pose(oracle):
    rng = np.random.default_rng(0)
    R = oracle.rodrigues(np.array([0.2, -0.3, 0.1]))
    t = np.array([50., -80., 2500.])
    uv = np.array([[100, 120], [500, 100], [320, 400], [200, 300]], np.float64)
    d = np.array([1500., 2500., 2000., 3000.])
    Xc = np.stack([(uv[:, 0] - 320) * d / 525, (uv[:, 1] - 240) * d / 525, d], 1)
    Y = (Xc - t) @ R
    ok, r, tt = oracle.solve_p3p(Y.astype(np.float32), uv.astype(np.float32))
    assert ok
    # float32 inputs: tolerance 1e-5 rad, 0.05 mm
    assert np.abs(oracle.rodrigues(r) - R).max() < 1e-5
    assert np.abs(tt - t).max() < 0.05


def test_diffmap_zero_for_gt_pose_on_clean_data(oracle, engine_mod):
    coords, pix, gt_cv, _ = _frame(engine_mod, rho=1.0, sigma=0.0)
    d = oracle.diff_map(coords, pix, gt_cv[:3], gt_cv[3:])
    # int16 rounding: <= 0.87 mm off -> <= 0.92 px at 500 mm depth
    assert d.max() < 0.95
    assert oracle.soft_inlier_score(d) > 0.1 * 1600 * 0.98


def test_diffmap_clamped_at_100(oracle, engine_mod):
    coords, pix, gt_cv, _ = _frame(engine_mod, rho=0.0)
    d = oracle.diff_map(coords, pix, gt_cv[:3], gt_cv[3:])
    assert d.max() == 100.0 and d.min() >= 0.0


def test_softmax_entropy(oracle):
    s = np.array([1.0, 2.0, 3.0, -1000.0])
    p = oracle.softmax(s)
    assert abs(p.sum() - 1) < 1e-15 and p[3] == 0.0
    e = np.exp(s[:3] - 3)
    assert np.abs(p[:3] - e / e.sum()).max() < 1e-15
    assert abs(oracle.entropy(np.full(8, 0.125)) - 3.0) < 1e-15
    assert oracle.entropy(np.array([1.0, 0.0])) == 0.0


def test_cv2our_roundtrip_and_loss(oracle):
    r = np.array([0.3, -0.2, 0.5]); t = np.array([100., -50., 2000.])
    R, tj = oracle.cv2our(r, t)
    assert abs(np.linalg.det(R) - 1) < 1e-12
    r2, t2 = oracle.our2cv(R, tj)
    assert np.abs(r2 - r).max() < 1e-12 and np.abs(t2 - t).max() < 1e-9
    loss, re, te = oracle.max_loss(R, tj, R, tj)
    assert loss < 1e-5 and te < 1e-9
    # 10 mm camera-centre shift -> tErr = 10, loss = max(rot, 1)
    R2, tj2 = oracle.cv2our(r, t + oracle.rodrigues(r) @ np.array([10., 0, 0]))
    loss, re, te = oracle.max_loss(R, tj, R2, tj2)
    assert abs(te - 10) < 1e-9 and abs(loss - 1.0) < 1e-9


def test_dloss_max_matches_finite_differences(oracle):
    rng = np.random.default_rng(3)
    for case in range(6):
        gt = np.concatenate([rng.uniform(-.5, .5, 3), rng.uniform(-1000, 1000, 3)])
        # translation-dominated and rotation-dominated cases
        est = gt + (np.concatenate([rng.normal(0, 1e-3, 3), rng.normal(0, 50, 3)]) if case % 2 == 0
                    else np.concatenate([rng.normal(0, 0.05, 3), rng.normal(0, 0.5, 3)]))
        jac = oracle.dloss_max(est, gt)

        def loss(e6):
            R1 = oracle.rodrigues(e6[:3]); R2 = oracle.rodrigues(gt[:3])
            rot = np.degrees(np.arccos(np.clip((np.trace(R1 @ R2.T) - 1) / 2, -1, 1)))
            tt = np.linalg.norm(R1.T @ (-e6[3:] / 10) - R2.T @ (-gt[3:] / 10))
            return rot, tt
        rot, tt = loss(est)
        fd = np.zeros(6)
        for i in range(6):
            h = 1e-6 if i < 3 else 1e-4
            a = est.copy(); a[i] += h
            b = est.copy(); b[i] -= h
            la, lb = loss(a), loss(b)
            k = 1 if tt > rot else 0
            fd[i] = (la[k] - lb[k]) / (2 * h)
        if tt > rot:
            # the reference omits the 1/10 of d(invT)/d(t) (quirk Q7): translation columns are 10x the true derivative
            assert np.abs(jac[3:] - 10 * fd[3:]).max() < 1e-4 * max(1, np.abs(jac[3:]).max())
            assert np.abs(jac[:3] - fd[:3]).max() < 1e-3 * max(1, np.abs(fd[:3]).max())
        else:
            assert np.abs(jac[:3] - fd[:3]).max() < 1e-3 * max(1, np.abs(fd[:3]).max())
            assert (jac[3:] == 0).all()


def test_dproject_matches_finite_differences(oracle):
    rng = np.random.default_rng(4)
    r = np.array([0.2, 0.1, -0.3]); t = np.array([30., -40., 2200.])
    R, tj = oracle.cv2our(r, t)
    for _ in range(10):
        obj = rng.uniform(-800, 800, 3).astype(np.float32)
        pt = rng.uniform(100, 500, 2).astype(np.float32)

        def err(o, Rm=R, tm=tj):
            e = Rm @ o.astype(np.float64) + tm
            px = -525 * e[0] / e[2] + 320; py = 525 * e[1] / e[2] + 240
            return np.hypot(pt[0] - px, pt[1] - py)
        if err(obj) > 100:
            assert (oracle.dproject_dobj(pt, obj, R, tj) == 0).all()
            continue
        j = oracle.dproject_dobj(pt, obj, R, tj)
        fd = np.array([(err(obj + h) - err(obj - h)) / 0.02 for h in np.eye(3, dtype=np.float32) * 0.01])
        assert np.abs(j - fd).max() < 1e-3 * max(1.0, np.abs(fd).max())
        j6 = oracle.dproject_dhyp(pt, obj, R, tj)
        rod = oracle.rodrigues_inv(R)
        fd6 = np.zeros(6)
        for i in range(6):
            h = 1e-6 if i < 3 else 1e-3
            ra, ta, rb, tb = rod.copy(), tj.copy(), rod.copy(), tj.copy()
            if i < 3:
                ra[i] += h; rb[i] -= h
            else:
                ta[i - 3] += h; tb[i - 3] -= h
            fd6[i] = (err(obj, oracle.rodrigues(ra), ta) - err(obj, oracle.rodrigues(rb), tb)) / (2 * h)
        assert np.abs(j6 - fd6).max() < 1e-4 * max(1.0, np.abs(fd6).max())


def test_kabsch_recovers_transform(oracle):
    rng = np.random.default_rng(7)
    R = oracle.rodrigues(np.array([0.4, -0.7, 0.2])); t = np.array([10., 20., -30.])
    a = rng.normal(0, 100, (12, 3))
    b = a @ R.T + t
    Ro, to = oracle.kabsch(a, b)
    assert np.abs(Ro - R).max() < 1e-12 and np.abs(to - t).max() < 1e-10


def test_forward_pipeline_recovers_pose(oracle, engine_mod):
    coords, pix, gt_cv, gt_jp = _frame(engine_mod)
    cfg = oracle.default_config()
    fw = oracle.forward(cfg, coords, pix, gt_jp[:9], gt_jp[9:])
    assert fw.status == 0
    assert abs(fw.sf.sum() - 1) < 1e-12
    assert fw.ref_steps_done == 8 and fw.n_perm_steps == 8
    assert fw.rot_err < 1.0 and fw.t_err < 30.0 and fw.correct == 1
    # sampled cells are distinct, poses reproject their own support within the threshold
    assert all(len(set(row)) == 4 for row in fw.img_idx.tolist())
    assert (np.diff(fw.cand_idx) > 0).all()
    for h in (0, 100, 255):
        d = oracle.diff_map(coords, pix, fw.hyp_rvec[h], fw.hyp_tvec[h])
        assert (d[fw.img_idx[h]] < 10).all()
        assert np.array_equal(d, fw.diffmaps[h])
    # inlier map: each step adds at most inlier_count, counts <= steps
    assert fw.inlier_map.max() <= 8 and fw.inlier_map.sum() <= 8 * 100


def test_refine_replay_equals_forward(oracle, engine_mod):
    coords, pix, gt_cv, gt_jp = _frame(engine_mod, f=2)
    cfg = oracle.default_config(seed=1305 + 2)
    fw = oracle.forward(cfg, coords, pix)
    rep = oracle.refine(cfg, fw.pixel_idxs, fw.n_perm_steps, coords, pix, fw.avg)
    assert np.abs(rep - oracle.jp6(fw.ref[:3], fw.ref[3:])).max() == 0.0


def test_streams_partition_like_libgomp(oracle, engine_mod):
    coords, pix, _, _ = _frame(engine_mod)
    a = oracle.forward(oracle.default_config(n_streams=1, ref_steps=0, n_hyps=10), coords, pix)
    b = oracle.forward(oracle.default_config(n_streams=3, ref_steps=0, n_hyps=10), coords, pix)
    # stream 0 of T=3 owns hypotheses [0,4) and draws the same candidates as the single stream
    assert np.array_equal(a.img_idx[:4], b.img_idx[:4])
    assert b.cand_idx[4] >= 0 and b.cand_idx[4] <= b.cand_idx[5]


def test_backward_runs_and_is_finite(oracle, engine_mod):
    coords, pix, gt_cv, gt_jp = _frame(engine_mod)
    cfg = oracle.default_config(n_hyps=16)
    fw = oracle.forward(cfg, coords, pix, gt_jp[:9], gt_jp[9:])
    bw = oracle.backward(cfg, coords, pix, gt_jp[:9], gt_jp[9:], fw)
    assert np.isfinite(bw.dloss_dobj).all()
    assert np.abs(bw.dloss_dobj).max() > 0
    # softmax-Jacobian gradients sum to zero (train_ransac_softam.cpp:361-376)
    assert abs(bw.score_grads.sum()) < 1e-9 * max(1.0, np.abs(bw.score_grads).max())


def test_upstream_restatement_known_answers(oracle):
    """Patch gather (cnn_softam.h:221-256 + train_obj.lua:117-124) and metres -> int16 mm (cnn_softam.h:262-268)."""
    O = oracle
    h, w = 60, 70
    frame = (np.arange(h * w * 3) % 251).astype(np.uint8).reshape(h, w, 3)
    pix = np.array([[21, 21], [w - 21, h - 21], [30, 25], [20, 30], [30, h - 20]])
    p = O.gather_patches(frame, pix)
    assert p.shape == (5, 3, 42, 42) and p.dtype == np.float32
    # patch(c, y, x) = frame(oy - 21 + y, ox - 21 + x)[c] - 127
    assert p[0, 0, 0, 0] == float(frame[0, 0, 0]) - 127 and p[0, 2, 41, 41] == float(frame[41, 41, 2]) - 127
    assert p[1, 1, 41, 41] == float(frame[h - 1, w - 1, 1]) - 127            # last legal centre reaches the last pixel
    assert p[2, 1, 3, 7] == float(frame[25 - 21 + 3, 30 - 21 + 7, 1]) - 127
    assert (p[3] == 0).all() and (p[4] == 0).all()                          # border cells are skipped by the reference
    c = O.coords_from_prediction(np.array([0.0005, 0.0015, 0.0025, -0.0015, 1.2344, 40.0, -40.0, np.nan, np.inf, 3e6], np.float32))
    assert c.tolist() == [0, 2, 2, -2, 1234, 32767, -32768, -32768, -32768, -32768]   # half-to-even; cvtss2si overflow -> INT_MIN
