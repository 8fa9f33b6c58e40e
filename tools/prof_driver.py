"""Minimal driver for ncu captures: a few full forward passes over NB frames (default 1024)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsac_b200 import engine as E
nb = int(os.environ.get("NB", "1024"))
reps = int(os.environ.get("REPS", "3"))
coords, pix, gt_cv, gt_jp = E.synth_frames(nb)
eng = E.Engine(max_frames=nb, write_diffmaps=int(os.environ.get("WRITE_DM", "1")))
dc = torch.from_numpy(coords).cuda(); dp = torch.from_numpy(pix).cuda(); dg = torch.from_numpy(gt_jp).cuda()
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
torch.cuda.synchronize()
print("done", eng.launches)
eng.close()
