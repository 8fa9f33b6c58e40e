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
