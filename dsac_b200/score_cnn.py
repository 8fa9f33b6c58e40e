"""Score-CNN plug-in behind the score seam.

The reference scores every hypothesis by pushing its 40x40 reprojection-error image through a Torch7 CNN
(`core/lua/train_score.lua:54-88`: 10 3x3 convolutions + 3 fully connected layers, input minus mean 45,
`core/lua/train_score_softam.lua:6,50-76`), reached through `forward(diffMaps, stateObj)`
(`core/lua_calls.h:284-300`).  This module is that network in PyTorch (library convolutions -- a dense conv net is
tensor-core work, unlike the engine's own kernels) wired to `dsac_set_score_hook`: it reads the engine's
DEVICE diffmap buffer in place and writes the H scores into the engine's device score buffer -- no host round trip
(the reference pushes 409 600 numbers through the Lua stack per frame, `core/lua_calls.h:89-105`).

The trained weights are not shipped with the reference (README.md:8), so the default is a deterministic random
initialisation; `load_state_dict` accepts converted weights.
"""
import torch
import torch.nn as nn

MEAN = 45.0   # train_score_softam.lua:6


def build_model(seed=0):
    g = torch.Generator().manual_seed(seed)
    def conv(i, o, s, p):
        return nn.Conv2d(i, o, 3, stride=s, padding=p)
    net = nn.Sequential(
        conv(1, 32, 1, 1), nn.ReLU(), conv(32, 32, 2, 1), nn.ReLU(),          # 40 -> 20
        conv(32, 64, 1, 1), nn.ReLU(), conv(64, 64, 2, 1), nn.ReLU(),         # 20 -> 10
        conv(64, 128, 1, 1), nn.ReLU(), conv(128, 128, 2, 1), nn.ReLU(),      # 10 -> 5
        conv(128, 256, 1, 1), nn.ReLU(), conv(256, 256, 2, 0), nn.ReLU(),     # 5 -> 2
        conv(256, 512, 1, 1), nn.ReLU(), conv(512, 512, 2, 1), nn.ReLU(),     # 2 -> 1
        nn.Flatten(), nn.Linear(512, 1024), nn.ReLU(), nn.Linear(1024, 1024), nn.ReLU(), nn.Linear(1024, 1),
    )
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 / max(1.0, p[0].numel() ** 0.5)) if p.dim() > 1
                    else torch.zeros(p.shape))
    return net.eval()


class _DevPtr:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


class ScoreCNN:
    """Callable for Engine.set_score_hook: fn(d_diffmaps, n, H, d_scores, stream) -> 0."""

    def __init__(self, device="cuda", seed=0, dtype=torch.float32, batch=8192):
        self.model = build_model(seed).to(device=device, dtype=dtype)
        self.device, self.dtype, self.batch = device, dtype, batch

    @torch.no_grad()
    def __call__(self, d_diffmaps, n, H, d_scores, stream):
        ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ext):
            dm = torch.as_tensor(_DevPtr(d_diffmaps, (n * H, 1, 40, 40), "<f4"), device=self.device)
            sc = torch.as_tensor(_DevPtr(d_scores, (n * H,), "<f8"), device=self.device)
            for lo in range(0, n * H, self.batch):
                x = (dm[lo:lo + self.batch] - MEAN).to(self.dtype)        # forward(): input[...]:add(-mean)
                sc[lo:lo + self.batch] = self.model(x).reshape(-1).double()
        return 0

    def backward(self, d_diffmaps, d_score_grads, n, H, d_diffmap_grads, stream):
        """Callable for Engine.set_score_backward_hook: the seam's adjoint `backward(maps, state, scoreOutputGradients,
        gradients)` (`core/lua_calls.h:312-341`, Lua side `core/lua/train_score_softam.lua:93-110`): push the H clamped
        output gradients (the engine clamps them to +-grad_clamp, `train_score_softam.lua:97`) back through the network to
        the H x 1600 inputs.  Reads / writes the engine's DEVICE buffers in place; parameter gradients accumulate in
        `self.model` (`.grad`) for an optimiser step by the caller, as the Lua side's `gradParams` do."""
        ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ext), torch.enable_grad():
            dm = torch.as_tensor(_DevPtr(d_diffmaps, (n * H, 1, 40, 40), "<f4"), device=self.device)
            sg = torch.as_tensor(_DevPtr(d_score_grads, (n * H,), "<f8"), device=self.device)
            out = torch.as_tensor(_DevPtr(d_diffmap_grads, (n * H, 1600), "<f8"), device=self.device)
            for lo in range(0, n * H, self.batch):
                x = (dm[lo:lo + self.batch] - MEAN).to(self.dtype).requires_grad_(True)
                y = self.model(x).reshape(-1)
                y.backward(sg[lo:lo + self.batch].to(self.dtype))
                out[lo:lo + self.batch] = x.grad.reshape(-1, 1600).double()
        return 0

