"""The product's device arithmetic (dsac_b200/csrc/*.cuh) compiled for the host and compared with
the oracle -- catches logic errors on the CPU box; the GPU parity tests proper are in test_gpu_*."""
import ctypes as C

import numpy as np


def _case(rng, noise, outlier):
    r = rng.uniform(-.5, .5, 3)
    t = np.array([rng.uniform(-300, 300), rng.uniform(-300, 300), rng.uniform(1500, 3000)])
    th = np.linalg.norm(r); k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    uv = np.stack([rng.integers(30, 610, 4), rng.integers(30, 450, 4)], 1).astype(np.float64)
    d = rng.uniform(500, 3500, 4)
    Xc = np.stack([(uv[:, 0] - 320) * d / 525, (uv[:, 1] - 240) * d / 525, d], 1)
    Y = np.round((Xc - t) @ R + rng.normal(0, noise, (4, 3))).astype(np.float32)
    if outlier:
        Y[rng.integers(0, 4)] = rng.uniform(-2000, 2000, 3).round()
    return Y, uv.astype(np.float32)


def test_minimal_set_hypothesis_matches_oracle(oracle, host_shim):
    rng = np.random.default_rng(11)
    n_acc = 0
    worst = 0.0
    for i in range(4000):
        Y, uv = _case(rng, [0, 5, 25][i % 3], i % 5 == 0)
        ok, ro, to = oracle.solve_p3p(Y, uv)
        acc = False
        if ok:
            p = oracle.project_points(Y.astype(np.float64), ro, to).astype(np.float32)
            d = uv - p
            acc = bool((np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2) < 10).all())
        rv, tv, fr = np.zeros(3), np.zeros(3), C.c_int(0)
        obj = np.ascontiguousarray(Y.reshape(-1)); img = np.ascontiguousarray(uv.reshape(-1))
        a2 = host_shim.shim_minimal_set(oracle._p(obj), oracle._p(img), C.c_double(525), C.c_double(320), C.c_double(240),
                                        10, oracle._p(rv), oracle._p(tv), C.byref(fr))
        assert bool(a2) == acc, i
        if acc:
            n_acc += 1
            worst = max(worst, np.abs(rv - ro).max(), np.abs(tv - to).max() / 1000)
    assert n_acc > 1000
    assert worst < 1e-9   # rad / m


def test_device_sampler_stream_matches_libstdcxx(oracle, host_shim):
    for seed, skip in ((1305, 6400), (1305, 0), (7, 13), (123456789, 1)):
        n = 700
        cells = np.zeros((n, 4), np.int32)
        used = np.zeros(n, np.uint32)
        host_shim.shim_candidates(C.c_uint32(seed), C.c_uint32(skip), n, oracle._p(cells), oracle._p(used))
        oc, od = oracle.candidates(seed, skip, n)
        assert np.array_equal(cells, oc[:, :, 1] * 40 + oc[:, :, 0])
        assert np.array_equal(used, od)   # no Lemire re-draws in these prefixes -> words == irand calls


def test_device_mt19937_matches_std(oracle, host_shim):
    out = np.zeros(2000, np.uint32)
    host_shim.shim_mt_raw(C.c_uint32(1305), 2000, oracle._p(out))
    assert np.array_equal(out, oracle.mt19937_raw(1305, 2000))


def test_block_parallel_mt19937_regeneration_matches_std(oracle, host_shim):
    """k_sample regenerates the whole 624-word state per barrier, thread t < 227 producing the words t, t+227, t+454 from
    old neighbours and its own results (mt_regenerate_words): same stream as std::mt19937, in any thread order."""
    for seed in (1305, 1306, 5489, 0xffffffff):
        out = np.zeros(5 * 624, np.uint32)
        host_shim.shim_mt_regenerated(C.c_uint32(seed), 5, oracle._p(out))
        assert np.array_equal(out, oracle.mt19937_raw(seed, 5 * 624))


def test_register_carried_regeneration_matches_std(oracle, host_shim):
    """mt_twist3 (the generator's and k1_spec's form of the regeneration: own words in registers, no branch) gives the same
    stream as std::mt19937."""
    for seed in (1305, 1306, 5489, 0xffffffff):
        out = np.zeros(7 * 624, np.uint32)
        host_shim.shim_mt_twist3(C.c_uint32(seed), 7, oracle._p(out))
        assert np.array_equal(out, oracle.mt19937_raw(seed, 7 * 624))


def test_pair_based_candidate_parser(host_shim):
    """cand_pairs_len (k_sample's boundary walk): number of (x, y) pairs a minimal set consumes = pairs drawn until four
    distinct cells are found (cnn_softam.h:1021-1039), incl. repeated cells, long runs of repeats and the end of the window."""
    rng = np.random.default_rng(21)

    def ref_len(p, sp, limit):
        seen = []
        q = sp
        while len(seen) < 4:
            if q >= limit:
                return -1
            if p[q] not in seen:
                seen.append(p[q])
            q += 1
        return q - sp

    for trial in range(300):
        n = 64
        # few distinct values -> many repeats (also more than 4 in a row)
        vals = rng.integers(0, [1600, 6, 3][trial % 3], n)
        p = ((vals // 40) << 8 | (vals % 40)).astype(np.uint16)
        for sp in range(0, n, 3):
            for limit in (n, min(n, sp + 5), min(n, sp + 9)):
                got = host_shim.shim_cand_pairs_len(p.ctypes.data_as(C.c_void_p), C.c_int(sp), C.c_int(limit))
                assert got == ref_len(list(p), sp, limit), (trial, sp, limit)


def test_stream_chunk_partition(host_shim):
    for H, T in ((256, 1), (256, 8), (10, 3), (7, 7), (64, 5)):
        covered = []
        for s in range(T):
            h0, cnt = C.c_int(0), C.c_int(0)
            host_shim.shim_stream_chunk(H, T, s, C.byref(h0), C.byref(cnt))
            covered += list(range(h0.value, h0.value + cnt.value))
            q, r = divmod(H, T)
            assert cnt.value == q + (1 if s < r else 0)
        assert covered == list(range(H))


def test_conservative_filter_never_rejects_an_accepted_candidate(host_shim, engine_mod):
    """K1's fp64 filter (p3p_quick_inline) may only say "certainly rejected" for candidates the exact
    P3P + reprojection check (cnn_softam.h:1041-1059) rejects; it should also flag only a few percent."""
    import ctypes as C
    E = engine_mod
    coords, pix, _, _ = E.synth_frames(4)
    tot = np.zeros(6, np.int64)
    for i in range(4):
        out = (C.c_longlong * 6)()
        host_shim.shim_filter_stats(coords[i].ctypes.data_as(C.c_void_p), pix[i].ctypes.data_as(C.c_void_p), C.c_uint32(1305 + i),
                                    C.c_uint32(6400 if i == 0 else 0), C.c_int(100000), C.c_double(525), C.c_double(320),
                                    C.c_double(240), C.c_int(10), out)
        tot += np.array(list(out))
    n, acc, flagged, _, missed, _ = [int(v) for v in tot]
    assert n == 400000 and acc > 3000
    assert missed == 0
    assert flagged < 0.03 * n


def _reproj_check(host_shim, rvec, tvec, coords, pix, want_err=False):
    n = len(coords)
    out = (C.c_longlong * 4)()
    err = np.zeros(n, np.float32) if want_err else None
    coords = np.ascontiguousarray(coords, np.int16); pix = np.ascontiguousarray(pix, np.int32)
    rvec = np.ascontiguousarray(rvec, np.float64); tvec = np.ascontiguousarray(tvec, np.float64)
    host_shim.shim_reproj_check(C.c_int(n), rvec.ctypes.data_as(C.c_void_p), tvec.ctypes.data_as(C.c_void_p),
                                coords.ctypes.data_as(C.c_void_p), pix.ctypes.data_as(C.c_void_p), C.c_double(525), C.c_double(320),
                                C.c_double(240), C.c_int(10), out, err.ctypes.data_as(C.c_void_p) if want_err else None)
    return [int(v) for v in out], err


def test_refinement_inlier_test_is_conservative(host_shim, engine_mod):
    """k_refine decides "reprojection error < threshold" in fp32 and falls back to the reference's exact arithmetic
    (getDiffMap, cnn_softam.h:319-362) when the fp32 error is within its error bound of the threshold.  A decided
    cell must always agree with the exact arithmetic; only a small fraction may be left undecided."""
    E = engine_mod
    rng = np.random.default_rng(5)
    coords, pix, gt_cv, _ = E.synth_frames(8)
    tot = np.zeros(4, np.int64)
    # (a) poses around the generating pose (what the refinement sees), incl. the exact generating pose
    for f in range(8):
        for k in range(40):
            s = 0.0 if k == 0 else 10.0 ** rng.uniform(-6, -1)
            rv = gt_cv[f, :3] + rng.normal(0, s, 3)
            tv = gt_cv[f, 3:] + rng.normal(0, 1000 * s, 3)
            o, _ = _reproj_check(host_shim, rv, tv, coords[f], pix[f])
            tot += o
    # (b) adversarial: for single cells, bisect one pose parameter until the exact error crosses the threshold
    #     (to float resolution), then test a cloud of poses around the crossing
    n_adv = 0
    for f in range(4):
        rv0, tv0 = gt_cv[f, :3].copy(), gt_cv[f, 3:].copy()
        _, e0 = _reproj_check(host_shim, rv0, tv0, coords[f], pix[f], True)
        cells = np.flatnonzero(e0 < 6.0)[:40]
        for c in cells:
            cc, pp = coords[f][c:c + 1], pix[f][c:c + 1]
            lo, hi = 0.0, 400.0     # shift of tvec[0] in mm: error grows from < 10 px to > 10 px
            if _reproj_check(host_shim, rv0, tv0 + [hi, 0, 0], cc, pp)[0][3] == 1:
                continue
            for _ in range(60):
                mid = 0.5 * (lo + hi)
                if _reproj_check(host_shim, rv0, tv0 + [mid, 0, 0], cc, pp)[0][3] == 1:
                    lo = mid
                else:
                    hi = mid
            for d in np.concatenate([[0.0], rng.normal(0, 1e-9, 8), rng.normal(0, 1e-6, 8), rng.normal(0, 1e-4, 8), rng.normal(0, 1e-2, 8)]):
                o, _ = _reproj_check(host_shim, rv0, tv0 + [lo + d, 0, 0], cc, pp)
                tot += o
                n_adv += 1
    # (c) extreme inputs: saturated coordinates, points at / behind the camera plane, huge translations
    cx = coords[0].copy()
    cx[::7] = 32767; cx[1::7] = -32768; cx[2::7] = 0
    for k in range(50):
        rv = rng.uniform(-3, 3, 3)
        tv = rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(0, 6)
        tot += _reproj_check(host_shim, rv, tv, cx, pix[0])[0]
    tot += _reproj_check(host_shim, np.zeros(3), np.zeros(3), cx, pix[0])[0]     # zero pose: z = Z, exactly 0 for some cells
    decided, undecided, wrong, inliers = [int(v) for v in tot]
    assert wrong == 0
    assert n_adv > 1000 and inliers > 100000
    assert undecided < 0.02 * (decided + undecided)


def test_lm_solve6_fast_matches_numpy(host_shim):
    """k_refine's damped 6x6 solve (Cholesky in registers, one rsqrt per column) against numpy; the singular case
    must go through the general (minimum-norm) path and stay finite."""
    rng = np.random.default_rng(3)
    iu = np.triu_indices(6)
    for i in range(300):
        J = rng.normal(size=(40, 6)) * 10.0 ** rng.uniform(-2, 3, 6)
        A = J.T @ J
        b = rng.normal(size=6) * np.sqrt(np.diag(A))
        lam = 10.0 ** rng.integers(-16, 3)
        S = np.concatenate([A[iu], b]).astype(np.float64)
        x = np.empty(6)
        host_shim.shim_lm_solve6_fast(S.ctypes.data_as(C.c_void_p), C.c_double(lam), x.ctypes.data_as(C.c_void_p))
        Ad = A.copy()
        Ad[np.diag_indices(6)] *= 1 + lam
        want = np.linalg.solve(Ad, b)
        assert np.abs(x - want).max() <= 1e-9 * np.abs(want).max() * max(1.0, np.linalg.cond(Ad) * 1e-7)
    A = np.zeros((6, 6)); A[0, 0] = 4.0
    S = np.concatenate([A[iu], [2.0, 0, 0, 0, 0, 0]])
    x = np.empty(6)
    host_shim.shim_lm_solve6_fast(S.ctypes.data_as(C.c_void_p), C.c_double(0.0), x.ctypes.data_as(C.c_void_p))
    assert np.allclose(x, [0.5, 0, 0, 0, 0, 0])


def test_rodrigues_jacobian_across_lanes_equals_sequential(host_shim):
    """rodrigues_jac_warp (one output per lane) = rodrigues_jac (sequential), bit for bit, incl. the zero-rotation branch."""
    rng = np.random.default_rng(4)
    for i in range(200):
        r = rng.uniform(-3, 3, 3) * (10.0 ** rng.uniform(-9, 0) if i % 3 else 1.0)
        if i == 0:
            r = np.zeros(3)
        R1, J1, R2, J2 = np.empty(9), np.empty(27), np.empty(9), np.empty(27)
        host_shim.shim_rodrigues_jac(r.ctypes.data_as(C.c_void_p), R1.ctypes.data_as(C.c_void_p), J1.ctypes.data_as(C.c_void_p))
        host_shim.shim_rodrigues_jac_lanes(r.ctypes.data_as(C.c_void_p), R2.ctypes.data_as(C.c_void_p), J2.ctypes.data_as(C.c_void_p))
        assert np.array_equal(R1, R2) and np.array_equal(J1, J2)


def test_lm_loop_of_the_refinement_matches_oracle(oracle, host_shim):
    """k_refine's Levenberg-Marquardt loop (one pass per trial vector, lm_advance state machine, register Cholesky),
    run sequentially on the host, against the oracle's cv::solvePnP(CV_ITERATIVE, useExtrinsicGuess) restatement
    (pinned to cv2 in test_oracle_golden): same iteration count, same pose."""
    rng = np.random.default_rng(8)
    dr, dt = [], []
    for i in range(200):
        n = int(rng.integers(50, 101))
        r = rng.uniform(-.5, .5, 3)
        t = np.array([rng.uniform(-300, 300), rng.uniform(-300, 300), rng.uniform(1500, 3000)])
        th = np.linalg.norm(r); k = r / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        uv = np.stack([rng.integers(21, 620, n), rng.integers(21, 460, n)], 1).astype(np.float64)
        d = rng.uniform(500, 3500, n)
        Xc = np.stack([(uv[:, 0] - 320) * d / 525, (uv[:, 1] - 240) * d / 525, d], 1)
        Y = np.round((Xc - t) @ R + rng.normal(0, [0, 5, 25][i % 3], (n, 3))).astype(np.float32)
        img = uv.astype(np.float32)
        r0 = r + rng.normal(0, 0.02, 3); t0 = t + rng.normal(0, 20, 3)
        ro, to, it_o = oracle.solve_pnp_iterative(Y, img, r0, t0)
        pose = np.concatenate([r0, t0]).astype(np.float64)
        passes = C.c_int(0)
        it = host_shim.shim_lm_refine(C.c_int(n), Y.ctypes.data_as(C.c_void_p), img.ctypes.data_as(C.c_void_p), C.c_double(525), C.c_double(320),
                                      C.c_double(240), pose.ctypes.data_as(C.c_void_p), C.byref(passes))
        assert it == it_o, (i, it, it_o)
        dr.append(np.abs(pose[:3] - ro).max()); dt.append(np.abs(pose[3:] - to).max())
    # the oracle solves the damped system with an SVD, the kernel with a Cholesky factorisation: rounding-level
    # differences (median 3e-17 rad); a noise-free problem whose final error comparison is decided by rounding
    # reaches 1.5e-10 rad / 7e-8 mm with either factorisation
    assert np.percentile(dr, 90) <= 1e-14 and np.percentile(dt, 90) <= 1e-11
    assert max(dr) <= 1e-9 and max(dt) <= 1e-6


def test_score_kernel_arithmetic_matches_the_exact_error_matrix(host_shim, engine_mod):
    """k_score's fp32 formulation (score_pair_error: one rsqrt with a floor, no per-pair guards; score_sigmoid_sum5: five
    sigmoids over one reciprocal), compiled for the host, against getDiffMap's exact arithmetic (cnn_softam.h:319-362)
    and the double-precision soft-inlier score: entries within the 2e-3 px contract, scores within 1e-5 relative + 5e-6."""
    E = engine_mod
    rng = np.random.default_rng(12)
    coords, pix, gt_cv, _ = E.synth_frames(6)

    def run(rv, tv, cc, pp):
        e32, ex = np.zeros(1600, np.float32), np.zeros(1600, np.float32)
        s, sx = C.c_double(0), C.c_double(0)
        cc = np.ascontiguousarray(cc, np.int16); pp = np.ascontiguousarray(pp, np.int32)
        rv = np.ascontiguousarray(rv, np.float64); tv = np.ascontiguousarray(tv, np.float64)
        host_shim.shim_score_hypothesis(rv.ctypes.data_as(C.c_void_p), tv.ctypes.data_as(C.c_void_p), cc.ctypes.data_as(C.c_void_p),
                                        pp.ctypes.data_as(C.c_void_p), C.c_double(525), C.c_double(320), C.c_double(240), C.c_double(10),
                                        C.c_double(0.1), C.c_double(0.5), e32.ctypes.data_as(C.c_void_p), ex.ctypes.data_as(C.c_void_p),
                                        C.byref(s), C.byref(sx))
        return e32, ex, s.value, sx.value

    worst = 0.0
    for f in range(6):
        for k in range(30):
            sc = 0.0 if k == 0 else 10.0 ** rng.uniform(-5, -0.5)
            rv = gt_cv[f, :3] + rng.normal(0, sc, 3)
            tv = gt_cv[f, 3:] + rng.normal(0, 1000 * sc, 3)
            e32, ex, s, sx = run(rv, tv, coords[f], pix[f])
            worst = max(worst, np.abs(e32 - ex).max())
            assert np.abs(e32 - ex).max() <= 2e-3
            assert abs(s - sx) <= 1e-5 * sx + 5e-6
    assert worst > 0        # (the two arithmetics do differ: fp32 vs double with float rounding of the projection)
    # value-encoded zero pose (R = I, t = 0): cells with Z = 0 lie in the camera plane (1/z := 1), cells at the origin
    # project onto the principal point; saturated coordinates must stay finite
    cz = coords[0].copy()
    cz[::3] = 0
    cz[1::11, 2] = 0
    cz[5::13] = 32767
    cz[7::17] = -32768
    e32, ex, s, sx = run(np.zeros(3), np.zeros(3), cz, pix[0])
    assert np.isfinite(e32).all() and np.abs(e32 - ex).max() <= 2e-3 and (e32[::3] < 100).any()
    assert abs(s - sx) <= 1e-5 * sx + 5e-6
    # a pose fitted exactly through a cell gives A = 0 there: the floor under the rsqrt must return exactly 0
    c1 = coords[0].copy(); p1 = pix[0].copy()
    c1[0] = (0, 0, 1000); p1[0] = (320, 240)
    e32, ex, _, _ = run(np.zeros(3), np.zeros(3), c1, p1)
    assert e32[0] == 0.0 and ex[0] == 0.0
