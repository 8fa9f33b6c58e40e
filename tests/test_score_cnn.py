"""the Score-CNN behind the seam.

CPU: the torch network of dsac_b200/score_cnn.py against an independent numpy restatement of the architecture the
reference builds in core/lua/train_score.lua:54-88 (ten 3x3 convolutions with the strides / paddings listed there, ReLU,
three fully connected layers, input minus mean 45, train_score_softam.lua:6) -- same weights, so this pins layer order,
strides, paddings and the flattening, which the self-consistency test on the GPU cannot.
GPU: ScoreCNN.backward (the seam's adjoint, lua_calls.h:312-341) against fp64 autograd of the same network."""
import numpy as np
import pytest

# (in, out, stride, pad) of the ten SpatialConvolution(i, o, 3, 3, s, s, p, p) layers, train_score.lua:56-75
CONVS = [(1, 32, 1, 1), (32, 32, 2, 1), (32, 64, 1, 1), (64, 64, 2, 1), (64, 128, 1, 1), (128, 128, 2, 1),
         (128, 256, 1, 1), (256, 256, 2, 0), (256, 512, 1, 1), (512, 512, 2, 1)]
FCS = [(512, 1024), (1024, 1024), (1024, 1)]   # train_score.lua:79-86


def conv3x3(x, w, b, stride, pad):
    """x [C, H, W], w [O, C, 3, 3] -> [O, H', W'] (cross-correlation, like nn.SpatialConvolution)."""
    c, h, wd = x.shape
    xp = np.zeros((c, h + 2 * pad, wd + 2 * pad)); xp[:, pad:pad + h, pad:pad + wd] = x
    ho, wo = (h + 2 * pad - 3) // stride + 1, (wd + 2 * pad - 3) // stride + 1
    out = np.zeros((w.shape[0], ho, wo))
    for i in range(ho):
        for j in range(wo):
            patch = xp[:, i * stride:i * stride + 3, j * stride:j * stride + 3]
            out[:, i, j] = np.tensordot(w, patch, axes=([1, 2, 3], [0, 1, 2])) + b
    return out


def numpy_score(params, diffmap):
    x = diffmap.reshape(1, 40, 40).astype(np.float64) - 45.0
    k = 0
    for (_i, _o, s, p) in CONVS:
        x = np.maximum(conv3x3(x, params[k], params[k + 1], s, p), 0.0); k += 2
    x = x.reshape(-1)
    for n, (_i, _o) in enumerate(FCS):
        x = params[k] @ x + params[k + 1]; k += 2
        if n < 2:
            x = np.maximum(x, 0.0)
    return float(x[0])


def test_torch_network_matches_numpy_restatement_of_the_reference_architecture():
    import torch
    from dsac_b200.score_cnn import build_model, MEAN
    assert MEAN == 45.0
    net = build_model(seed=5).double()
    params = [p.detach().numpy() for p in net.parameters()]
    shapes = [tuple(p.shape) for p in params]
    want = []
    for (i, o, _s, _p) in CONVS:
        want += [(o, i, 3, 3), (o,)]
    for (i, o) in FCS:
        want += [(o, i), (o,)]
    assert shapes == want
    # biases are zero in the default initialisation: give them values so that the test sees them
    rng = np.random.default_rng(0)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.copy_(torch.from_numpy(rng.normal(0, 0.05, p.shape)))
    params = [p.detach().numpy() for p in net.parameters()]
    maps = rng.uniform(0, 100, (3, 40, 40))          # reprojection errors are clamped to [0, 100] px (cnn_softam.h:355-358)
    with torch.no_grad():
        got = net(torch.from_numpy(maps).reshape(3, 1, 40, 40) - MEAN).reshape(-1).numpy()
    for m in range(3):
        ref = numpy_score(params, maps[m])
        assert abs(got[m] - ref) <= 1e-9 * max(1.0, abs(ref)), (m, got[m], ref)


@pytest.mark.gpu
def test_score_cnn_backward_behind_the_seam(engine_mod):
    """The CNN's adjoint through dsac_set_score_backward_hook: what the hook writes (fp32 network) equals fp64 autograd of
    the same network on the same device diffmaps fed with the same (clamped) output gradients, and the engine's gradient
    w.r.t. the scene coordinates is finite and not the closed-form one."""
    import copy
    import torch
    from dsac_b200.score_cnn import ScoreCNN, MEAN
    E = engine_mod
    nf, H = 2, 32
    coords, pix, gt_cv, gt_jp = E.synth_frames(nf)
    torch.backends.cudnn.allow_tf32 = False          # the comparison below is fp32 vs fp64, not tf32 vs fp64
    torch.backends.cuda.matmul.allow_tf32 = False
    cnn = ScoreCNN(seed=3)
    ref_net = copy.deepcopy(cnn.model).double()
    seen = {}

    class _W:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

    def bwd(dm, sg, n, Hh, out, stream):
        rc = cnn.backward(dm, sg, n, Hh, out, stream)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            d = torch.as_tensor(_W(dm, (n * Hh, 1, 40, 40), "<f4"), device="cuda")
            g = torch.as_tensor(_W(sg, (n * Hh,), "<f8"), device="cuda")
            o = torch.as_tensor(_W(out, (n * Hh, 1600), "<f8"), device="cuda")
            x = (d.double() - MEAN).requires_grad_(True)
            ref_net(x).reshape(-1).backward(g.clone())
            seen["got"], seen["want"], seen["gmax"] = o.clone().cpu().numpy(), x.grad.reshape(-1, 1600).cpu().numpy(), float(g.abs().max())
        return rc

    eng = E.Engine(max_frames=nf, n_hyps=H)
    eng.set_score_hook(cnn)
    eng.set_score_backward_hook(bwd)
    eng.forward(coords, pix, gt_jp)
    got = eng.backward(coords, pix, gt_jp, full=False)
    eng.close()
    plain = E.Engine(max_frames=nf, n_hyps=H)
    plain.forward(coords, pix, gt_jp)
    closed = plain.backward(coords, pix, gt_jp, full=False)
    plain.close()
    assert seen["gmax"] <= 0.1 + 1e-15, seen["gmax"]                    # clamped output gradients (train_score_softam.lua:97)
    scale = np.abs(seen["want"]).max()
    dev = np.abs(seen["got"] - seen["want"]).max()
    # fp32 network vs fp64 autograd: measured 1.2e-2 of the largest entry (a ReLU whose pre-activation is ~0 switches with the
    # precision; the gradients of this random initialisation are ~1e-8)
    assert scale > 0 and dev <= 5e-2 * scale, (dev, scale)
    assert np.isfinite(got.dloss_dobj).all() and np.abs(got.dloss_dobj).max() > 0
    assert np.abs(got.dloss_dobj - closed.dloss_dobj).max() > 0         # it is the CNN's gradient, not the closed form's
    assert any(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in cnn.model.parameters())   # gradParams accumulated
