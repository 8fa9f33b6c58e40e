"""C-ABI library: loads on a CPU-only box, exports every symbol include/dsac_b200.h declares,
host helpers work, and the engine refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(engine_mod):
    hdr = open(os.path.join(ROOT, "include", "dsac_b200.h")).read()
    declared = set(re.findall(r"\b(dsac_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dsac_score_hook"}
    assert declared == set(engine_mod.EXPORTS), declared ^ set(engine_mod.EXPORTS)
    lib = engine_mod.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.dsac_version()


def test_default_config_matches_reference_defaults(engine_mod):
    c = engine_mod.default_config()
    # properties.cpp:39-83 and 7scenes default.config
    assert (c.focal, c.cx, c.cy) == (525.0, 320.0, 240.0)
    assert (c.n_hyps, c.thr2d, c.inlier_count, c.ref_steps) == (256, 10, 100, 8)
    assert abs(c.sub_sample - 0.01) < 1e-15 and c.seed == 1305 and c.stream_skip == 6400


def test_config_struct_layout_matches_header(engine_mod):
    # sizeof(dsac_config) as the C compiler lays it out: 3*8 + 4*4 + 3*8 + 4*4 + 4(+4 pad) + 8 + 4*4 = 112
    assert C.sizeof(engine_mod.Config) == 112
    assert C.sizeof(engine_mod.ForwardOut) == 18 * 8


def test_engine_create_fails_loudly_without_gpu(engine_mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError) as ei:
        engine_mod.Engine()
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_reference_oracle():
    """The product path must never route through the oracle."""
    for d, _, files in os.walk(os.path.join(ROOT, "dsac_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "dsac_oracle" not in txt, f
    for d, _, files in os.walk(os.path.join(ROOT, "apps")):
        for f in files:
            if not f.endswith((".cpp", ".h", "Makefile")):
                continue
            txt = open(os.path.join(d, f)).read()
            assert "dsac_oracle" not in txt, f


def test_synth_frames_deterministic_and_sharding_invariant(engine_mod, oracle):
    a = engine_mod.synth_frames(6)
    b = engine_mod.synth_frames(6)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    c = engine_mod.synth_frames(3, frame0=3)
    for x, y in zip(a, c):
        assert np.array_equal(x[3:], y)
    coords, pix, gt_cv, gt_jp = a
    assert coords.dtype == np.int16 and pix.dtype == np.int32
    # grid of frame g comes from mt19937(seed + g): equals the oracle's stochasticSubSample restatement
    for g in range(3):
        assert np.array_equal(pix[g], oracle.stochastic_subsample(1305 + g))
    assert pix[:, :, 0].min() >= 21 and pix[:, :, 0].max() <= 640 - 21
    # gt_jp is cv2our(gt_cv)
    R, t = oracle.cv2our(gt_cv[0, :3], gt_cv[0, 3:])
    assert np.abs(gt_jp[0, :9].reshape(3, 3) - R).max() < 1e-14 and np.abs(gt_jp[0, 9:] - t).max() < 1e-12
    # about half the points are inliers of the GT pose
    d = oracle.diff_map(coords[0], pix[0], gt_cv[0, :3], gt_cv[0, 3:])
    frac = (d < 30).mean()
    assert 0.35 < frac < 0.65


def test_stochastic_subsample_semantics(engine_mod):
    pix = engine_mod.stochastic_subsample(1305).reshape(40, 40, 2)
    # cell (sy, sx): x in [21 + 14.95 sx, +14.95), y in [21 + 10.95 sy, +10.95)  (cnn_softam.h:288-300)
    for sy in (0, 7, 39):
        for sx in (0, 13, 39):
            x, y = pix[sy, sx]
            assert 21 + 14.95 * sx - 1 <= x <= 21 + 14.95 * (sx + 1) + 1
            assert 21 + 10.95 * sy - 1 <= y <= 21 + 10.95 * (sy + 1) + 1
