"""Generates tests/golden/*.npz.

The reference (cvlab-dresden/DSAC) ships no tests or golden vectors and its
numeric kernels are OpenCV calls (OpenCV 2.4, not vendored, no pinned version).  These
fixtures pin the OpenCV boundary of the oracle against the only OpenCV available here,
Python cv2 (version recorded in the file), and pin the libstdc++ RNG contract
(std::mt19937 / uniform_int_distribution / std::shuffle of this image's libstdc++ 13).

    python tests/golden/make_golden.py      # needs cv2; run in the build container only
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

K = np.array([[525, 0, 320], [0, 525, 240], [0, 0, 1]], np.float64)


def rand_case(rng, n, noise):
    import cv2
    r = rng.uniform(-.5, .5, 3)
    t = np.array([rng.uniform(-300, 300), rng.uniform(-300, 300), rng.uniform(1500, 3000)])
    R = cv2.Rodrigues(r)[0]
    uv = np.stack([rng.integers(30, 610, n), rng.integers(30, 450, n)], 1).astype(np.float64)
    d = rng.uniform(500, 3500, n)
    Xc = np.stack([(uv[:, 0] - 320) * d / 525, (uv[:, 1] - 240) * d / 525, d], 1)
    Y = np.round((Xc - t) @ R + rng.normal(0, noise, (n, 3))).astype(np.float32)
    return Y, uv.astype(np.float32), r, t


def main():
    import cv2
    rng = np.random.default_rng(20170721)
    out = {"cv2_version": np.array(cv2.__version__)}
    # Rodrigues (+ Jacobian) both ways
    rv = rng.uniform(-2.5, 2.5, (64, 3))
    rv[0] = 0
    rv[1] = [1e-9, 0, 0]
    rv[2] = [np.pi, 0, 0]
    rv[3] = [0, np.pi - 1e-7, 0]
    Rm, Jm, rback = [], [], []
    for r in rv:
        R, J = cv2.Rodrigues(r.reshape(3, 1))
        Rm.append(R)
        Jm.append(J)
        rback.append(cv2.Rodrigues(R)[0].ravel())
    out.update(rod_r=rv, rod_R=np.array(Rm), rod_J=np.array(Jm), rod_back=np.array(rback))
    # projectPoints with Jacobians
    X = rng.uniform(-2000, 2000, (40, 3))
    pr = rng.uniform(-.5, .5, 3)
    pt = np.array([10., -20., 2500.])
    uv, jac = cv2.projectPoints(X, pr, pt, K, None)
    out.update(proj_X=X, proj_r=pr, proj_t=pt, proj_uv=uv.reshape(-1, 2), proj_jac=jac[:, :6])
    # P3P (4 points)
    objs, imgs, oks, rs, ts = [], [], [], [], []
    for i in range(400):
        Y, uvf, _, _ = rand_case(rng, 4, [0.0, 5.0, 25.0][i % 3])
        if i % 7 == 0:
            Y[rng.integers(0, 4)] = rng.uniform(-2000, 2000, 3).round()
        ok, r, t = cv2.solvePnP(Y.reshape(-1, 1, 3), uvf.reshape(-1, 1, 2), K, None, flags=cv2.SOLVEPNP_P3P)
        ok = bool(ok) and not (np.isnan(r).any() or np.isnan(t).any())
        objs.append(Y); imgs.append(uvf); oks.append(ok)
        rs.append(r.ravel() if ok else np.zeros(3)); ts.append(t.ravel() if ok else np.zeros(3))
    out.update(p3p_obj=np.array(objs), p3p_img=np.array(imgs), p3p_ok=np.array(oks), p3p_r=np.array(rs), p3p_t=np.array(ts))
    # iterative PnP with extrinsic guess
    lobj, limg, lr0, lt0, lr, lt, ln = [], [], [], [], [], [], []
    for i in range(40):
        n = int(rng.integers(50, 101))
        Y, uvf, r, t = rand_case(rng, n, 10.0)
        r0 = r + rng.normal(0, 0.02, 3)
        t0 = t + rng.normal(0, 20, 3)
        ok, r2, t2 = cv2.solvePnP(Y.reshape(-1, 1, 3), uvf.reshape(-1, 1, 2), K, None, rvec=r0.reshape(3, 1).copy(),
                                  tvec=t0.reshape(3, 1).copy(), useExtrinsicGuess=True, flags=cv2.SOLVEPNP_ITERATIVE)
        Yp = np.zeros((100, 3), np.float32); Yp[:n] = Y
        up = np.zeros((100, 2), np.float32); up[:n] = uvf
        lobj.append(Yp); limg.append(up); ln.append(n); lr0.append(r0); lt0.append(t0); lr.append(r2.ravel()); lt.append(t2.ravel())
    out.update(lm_obj=np.array(lobj), lm_img=np.array(limg), lm_n=np.array(ln), lm_r0=np.array(lr0), lm_t0=np.array(lt0),
               lm_r=np.array(lr), lm_t=np.array(lt))
    # SVD 3x3 (for Kabsch): singular values only are unique
    A = rng.normal(0, 1, (16, 3, 3))
    out.update(svd_A=A, svd_w=np.array([cv2.SVDecomp(a)[0].ravel() for a in A]))
    np.savez_compressed(os.path.join(HERE, "cv2_golden.npz"), **out)

    # RNG contract of this image's libstdc++ (through the oracle, which calls it directly)
    from oracle import oracle as O
    cells, draws = O.candidates(1305, 6400, 512)
    np.savez_compressed(os.path.join(HERE, "rng_golden.npz"),
                        mt_raw_1305=O.mt19937_raw(1305, 1300), mt_raw_5489=O.mt19937_raw(5489, 16),
                        cand_cells=cells.astype(np.int16), cand_draws=draws.astype(np.uint8),
                        subsample_1305=O.stochastic_subsample(1305).astype(np.int16),
                        perm=O.refine_permutations(8).astype(np.int16))
    print("wrote golden fixtures with cv2", cv2.__version__)


if __name__ == "__main__":
    main()
