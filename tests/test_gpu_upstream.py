"""the upstream step on the device -- patch gather (cnn_softam.h:221-256 +
lua/train_obj.lua:117-124 + lua_calls.h:65-82) and the metres -> int16 mm conversion (cnn_softam.h:262-268) -- against
their numpy restatement in oracle/oracle.py.  Byte / integer work: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(n, w=640, h=480, seed=5):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("shared", [0, 1])
def test_gather_patches_bit_exact(engine_mod, oracle, shared):
    import torch
    E, O = engine_mod, oracle
    n = 3
    frames = _frames(n)
    pix = np.stack([E.stochastic_subsample(1305 + f) for f in range(1 if shared else n)]).astype(np.int32)   # [n or 1][1600][2]
    eng = E.Engine(max_frames=n, n_hyps=8)
    d_frames = torch.from_numpy(frames).cuda(); d_pix = torch.from_numpy(pix).cuda()
    d_patches = torch.full((n, E.N, 3, 42, 42), -1.0, dtype=torch.float32, device="cuda")
    d_status = torch.ones(n, dtype=torch.int32, device="cuda")
    eng.gather_patches_device(n, d_frames.data_ptr(), 640, 480, d_pix.data_ptr(), shared, d_patches.data_ptr(), mean=127.0,
                              d_status=d_status.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    got = d_patches.cpu().numpy()
    for f in range(n):
        want = O.gather_patches(frames[f], pix[0 if shared else f])
        assert np.array_equal(got[f], want)
    assert (d_status.cpu().numpy() == 0).all()
    assert got.min() >= -127.0 and got.max() <= 128.0


def test_gather_patches_edges(engine_mod, oracle):
    """Cells on the frame border (skipped by the reference) come back as zero patches + a status flag; the extreme legal
    centres (21 and width-21 / height-21) are gathered; small frames are handled."""
    import torch
    E, O = engine_mod, oracle
    w, h = 100, 90
    frames = _frames(2, w, h, seed=9)
    pix = np.zeros((2, E.N, 2), np.int32)
    rng = np.random.default_rng(1)
    pix[..., 0] = rng.integers(21, w - 21 + 1, size=(2, E.N)); pix[..., 1] = rng.integers(21, h - 21 + 1, size=(2, E.N))
    pix[0, 0] = (21, 21); pix[0, 1] = (w - 21, h - 21)            # extreme legal centres
    pix[1, 5] = (20, 40); pix[1, 6] = (w - 20, 40); pix[1, 7] = (50, h - 20); pix[1, 8] = (50, 3)   # border cells
    eng = E.Engine(max_frames=2, n_hyps=8)
    d_frames = torch.from_numpy(frames).cuda(); d_pix = torch.from_numpy(pix).cuda()
    d_patches = torch.full((2, E.N, 3, 42, 42), 7.0, dtype=torch.float32, device="cuda")
    d_status = torch.zeros(2, dtype=torch.int32, device="cuda")
    eng.gather_patches_device(2, d_frames.data_ptr(), w, h, d_pix.data_ptr(), 0, d_patches.data_ptr(), d_status=d_status.data_ptr())
    torch.cuda.synchronize()
    got = d_patches.cpu().numpy()
    for f in range(2):
        assert np.array_equal(got[f], O.gather_patches(frames[f], pix[f]))
    assert (got[1, 5:9] == 0).all()
    st = d_status.cpu().numpy()
    assert st[0] == 0 and st[1] == 4      # DSAC_ST_BORDER_PATCH
    with pytest.raises(RuntimeError):
        eng.gather_patches_device(1, d_frames.data_ptr(), 30, 30, d_pix.data_ptr(), 0, d_patches.data_ptr())


def test_coords_from_prediction_bit_exact(engine_mod, oracle):
    import torch
    E, O = engine_mod, oracle
    n = 2
    rng = np.random.default_rng(3)
    pred = (rng.standard_normal((n, E.N, 3)) * 3).astype(np.float32)
    flat = pred.reshape(-1)
    flat[:12] = [0.0005, 0.0015, 0.0025, -0.0005, -0.0015, 32.767, 32.7675, 40.0, -32.768, -33.0, 0.0, -0.0]   # ties, saturation
    flat[12:16] = [np.nan, np.inf, -np.inf, 3e6]
    eng = E.Engine(max_frames=n, n_hyps=8)
    d_pred = torch.from_numpy(pred).cuda()
    d_coords = torch.zeros((n, E.N, 3), dtype=torch.int16, device="cuda")
    eng.coords_from_prediction_device(n, d_pred.data_ptr(), d_coords.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d_coords.cpu().numpy(), O.coords_from_prediction(pred))


def test_frame_to_hypotheses_stays_on_the_device(engine_mod):
    """BGR frames -> patches -> (library) coordinate CNN -> int16 grid -> hypothesis engine, all on device pointers."""
    import torch
    from dsac_b200.coord_cnn import CoordPipeline
    E = engine_mod
    n = 2
    frames = torch.from_numpy(_frames(n)).cuda()
    pix = torch.from_numpy(np.stack([E.stochastic_subsample(1305 + f) for f in range(n)]).astype(np.int32)).cuda()
    eng = E.Engine(max_frames=n, n_hyps=16, max_candidates=20000)
    pipe = CoordPipeline(eng, seed=1, chunk=1)
    coords = pipe(frames, pix)
    assert coords.dtype == torch.int16 and tuple(coords.shape) == (n, E.N, 3)
    eng.forward_device(n, coords.data_ptr(), pix.data_ptr(), 0, None, 0, torch.cuda.current_stream().cuda_stream)
    res = eng.fetch(n)
    assert np.isfinite(res.scores).all() and res.img_idx.shape == (n, 16, 4)
