"""k_gather_patches in isolation: 64 frames x 1600 patches, L2 flushed between launches; checks against numpy first."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsac_b200 import engine as E
from oracle import oracle as O
nb = 64
rng = np.random.default_rng(7)
frames = rng.integers(0, 256, size=(nb, 480, 640, 3), dtype=np.uint8)
pix = np.stack([E.stochastic_subsample(1305 + f) for f in range(nb)]).astype(np.int32)
fr = torch.from_numpy(frames).cuda(); px = torch.from_numpy(pix).cuda()
patches = torch.empty((nb, E.N, 3, 42, 42), dtype=torch.float32, device="cuda")
eng = E.Engine(max_frames=1, n_hyps=8)
st = torch.cuda.current_stream().cuda_stream
eng.gather_patches_device(nb, fr.data_ptr(), 640, 480, px.data_ptr(), 0, patches.data_ptr(), stream=st)
torch.cuda.synchronize()
assert np.array_equal(patches[3].cpu().numpy(), O.gather_patches(frames[3], pix[3]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(12):
    flush.zero_()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); eng.gather_patches_device(nb, fr.data_ptr(), 640, 480, px.data_ptr(), 0, patches.data_ptr(), stream=st); b.record()
    torch.cuda.synchronize()
    if i >= 2: ts.append(a.elapsed_time(b))
ms = sum(ts) / len(ts)
by = nb * (E.N * 3 * 42 * 42 * 4 + 480 * 640 * 3 + E.N * 8)
print("k_gather_patches %.3f ms  %.0f GB/s" % (ms, by / ms / 1e6))
# reference point: a plain device memset of the same size
t2 = []
for i in range(6):
    flush.zero_()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); patches.zero_(); b.record(); torch.cuda.synchronize()
    if i >= 1: t2.append(a.elapsed_time(b))
print("memset of the patch buffer %.3f ms  %.0f GB/s" % (sum(t2) / len(t2), patches.numel() * 4 / (sum(t2) / len(t2)) / 1e6))
