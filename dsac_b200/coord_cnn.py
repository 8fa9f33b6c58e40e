"""the coordinate CNN of the reference (core/lua/train_obj.lua:47-98: ten 3x3
convolutions 64-64-128-128-256-256-256-512-512-512, three FC layers 4096-4096-3, ReLU) as a torch/cuDNN plug-in
between the two upstream kernels of the engine, so a frame goes BGR image -> patches -> scene coordinates -> hypotheses
without leaving the GPU:

    dsac_gather_patches_device   (CUDA, dsac_b200/csrc/upstream.cuh)   frames + sampling grid -> [n*1600][3][42][42]
    CoordCNN                      (torch / cuDNN: LIBRARY code, like the Score CNN of score_cnn.py)
    dsac_coords_from_prediction_device (CUDA)                          metres -> int16 mm grid the engine consumes

The network is random-initialised here (no trained weights offline); the convolutional trunk is out of this repo's scope
 -- only the data path around it is product code."""
import torch
import torch.nn as nn

from .engine import N

PATCH = 42
MEAN = 127.0


def build_model():
    def conv(i, o, s, p):
        return [nn.Conv2d(i, o, 3, s, p), nn.ReLU(inplace=True)]
    layers = (conv(3, 64, 1, 0) + conv(64, 64, 2, 1) + conv(64, 128, 1, 1) + conv(128, 128, 2, 1) + conv(128, 256, 1, 1) +
              conv(256, 256, 1, 1) + conv(256, 256, 2, 1) + conv(256, 512, 1, 1) + conv(512, 512, 1, 1) + conv(512, 512, 2, 0))
    return nn.Sequential(*layers, nn.Flatten(), nn.Linear(2 * 2 * 512, 4096), nn.ReLU(inplace=True), nn.Linear(4096, 4096),
                         nn.ReLU(inplace=True), nn.Linear(4096, 3))


class CoordPipeline:
    """frames (uint8 BGR, on the device) -> int16 scene-coordinate grids (on the device), in chunks of `chunk` frames."""

    def __init__(self, engine, seed=0, chunk=8, batch=1600, device="cuda"):
        torch.manual_seed(seed)
        self.engine = engine
        self.model = build_model().to(device).eval()
        self.chunk, self.batch, self.device = chunk, batch, device
        self.patches = torch.empty((chunk * N, 3, PATCH, PATCH), dtype=torch.float32, device=device)

    @torch.no_grad()
    def __call__(self, frames, pix):
        """frames: uint8 tensor [n][H][W][3] on the device; pix: int32 tensor [n][1600][2] (or [1600][2], shared)."""
        n, height, width = frames.shape[0], frames.shape[1], frames.shape[2]
        shared = 1 if pix.numel() == N * 2 else 0
        coords = torch.empty((n, N, 3), dtype=torch.int16, device=self.device)
        st = torch.cuda.current_stream().cuda_stream
        for f0 in range(0, n, self.chunk):
            m = min(self.chunk, n - f0)
            px = pix if shared else pix[f0:f0 + m]
            self.engine.gather_patches_device(m, frames[f0:f0 + m].data_ptr(), width, height, px.data_ptr(), shared,
                                              self.patches.data_ptr(), mean=MEAN, stream=st)
            flat = self.patches[:m * N]
            pred = torch.cat([self.model(flat[i:i + self.batch]) for i in range(0, m * N, self.batch)]).float().contiguous()
            self.engine.coords_from_prediction_device(m, pred.data_ptr(), coords[f0:f0 + m].data_ptr(), stream=st)
        return coords
