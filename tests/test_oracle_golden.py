"""Oracle vs committed golden vectors (cv2 4.13 for the OpenCV boundary, libstdc++ 13 for the RNG).

The reference has no tests/golden vectors of its own; these fixtures are the
pin for the oracle.  Tolerances are stated per check.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "cv2_golden.npz"))


@pytest.fixture(scope="module")
def rg():
    return np.load(os.path.join(GOLDEN, "rng_golden.npz"))


def test_rodrigues_matches_cv2(oracle, g):
    for r, R, J, rb in zip(g["rod_r"], g["rod_R"], g["rod_J"], g["rod_back"]):
        Ro, Jo = oracle.rodrigues(r, jac=True)
        assert np.abs(Ro - R).max() <= 1e-14
        assert np.abs(Jo - J).max() <= 1e-13
        # matrix -> vector: tolerance 1e-9 (acos near pi amplifies rounding)
        assert np.abs(oracle.rodrigues_inv(R) - rb).max() <= 1e-9


def test_project_points_matches_cv2(oracle, g):
    uv, dr, dt = oracle.project_points(g["proj_X"], g["proj_r"], g["proj_t"], jac=True)
    assert np.abs(uv - g["proj_uv"]).max() <= 1e-10      # pixels
    jac = g["proj_jac"]
    assert np.abs(dr - jac[:, 0:3]).max() <= 1e-8 * max(1.0, np.abs(jac[:, 0:3]).max())
    assert np.abs(dt - jac[:, 3:6]).max() <= 1e-10


def test_p3p_matches_cv2(oracle, g):
    n_ok = 0
    for obj, img, ok, r, t in zip(g["p3p_obj"], g["p3p_img"], g["p3p_ok"], g["p3p_r"], g["p3p_t"]):
        oko, ro, to = oracle.solve_p3p(obj, img)
        assert oko == bool(ok)
        if not ok:
            assert (ro == 0).all() and (to == 0).all()   # safeSolvePnP zero pose, cnn_softam.h:66-71
            continue
        n_ok += 1
        # pose tolerance: 1e-8 rad / 1e-5 mm (cv2 4.13's P3P is not bit-identical to ours)
        assert np.abs(ro - r).max() <= 1e-8
        assert np.abs(to - t).max() <= 1e-5
    assert n_ok > 300


def test_iterative_pnp_matches_cv2(oracle, g):
    for obj, img, n, r0, t0, r, t in zip(g["lm_obj"], g["lm_img"], g["lm_n"], g["lm_r0"], g["lm_t0"], g["lm_r"], g["lm_t"]):
        ro, to, it = oracle.solve_pnp_iterative(obj[:n], img[:n], r0, t0)
        assert 1 <= it <= 20
        assert np.abs(ro - r).max() <= 1e-10              # rad
        assert np.abs(to - t).max() <= 1e-7               # mm


def test_svd3_matches_cv2(oracle, g):
    for A, w in zip(g["svd_A"], g["svd_w"]):
        U, wo, Vt = oracle.svd3(A)
        assert np.abs(wo - w).max() <= 1e-12
        assert np.abs(U @ np.diag(wo) @ Vt - A).max() <= 1e-12
        assert np.abs(U.T @ U - np.eye(3)).max() <= 1e-12


def test_live_cv2_if_present(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    K = np.array([[525, 0, 320], [0, 525, 240], [0, 0, 1]], np.float64)
    for _ in range(50):
        X = rng.uniform(-2000, 2000, (10, 3))
        r = rng.uniform(-1, 1, 3)
        t = np.array([0., 0., 3000.])
        assert np.abs(oracle.project_points(X, r, t) - cv2.projectPoints(X, r, t, K, None)[0].reshape(-1, 2)).max() <= 1e-10


def test_rng_contract_mt19937(oracle, rg):
    assert (oracle.mt19937_raw(1305, 1300) == rg["mt_raw_1305"]).all()
    assert (oracle.mt19937_raw(5489, 16) == rg["mt_raw_5489"]).all()
    # first output of a default-seeded mt19937 is the textbook value
    assert int(rg["mt_raw_5489"][0]) == 3499211612


def test_rng_contract_candidates_and_perm(oracle, rg):
    cells, draws = oracle.candidates(1305, 6400, 512)
    assert (cells == rg["cand_cells"]).all()
    assert (draws == rg["cand_draws"]).all()
    assert (oracle.stochastic_subsample(1305) == rg["subsample_1305"]).all()
    perm = oracle.refine_permutations(8)
    assert (perm == rg["perm"]).all()
    for s in range(8):
        assert sorted(perm[s].tolist()) == list(range(1600))


def test_candidates_are_lemire_of_raw_stream(oracle):
    """libstdc++'s uniform_int_distribution<int>(0,39) on mt19937 == (word*40)>>32 with rejection below 16."""
    raw = oracle.mt19937_raw(1305, 6400 + 4096).astype(np.uint64)
    cells, draws = oracle.candidates(1305, 6400, 256)
    pos = 6400
    for k in range(256):
        got = []
        while len(got) < 4:
            vals = []
            for _ in range(2):
                while True:
                    prod = raw[pos] * np.uint64(40)
                    pos += 1
                    if int(prod & np.uint64(0xffffffff)) >= 16:
                        break
                vals.append(int(prod >> np.uint64(32)))
            c = (vals[0], vals[1])
            if c in got:
                continue
            got.append(c)
        assert [list(c) for c in got] == cells[k].tolist()
