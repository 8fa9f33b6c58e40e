"""K2 (k_score) in isolation: time with and without the diffmap store (WRITE_DM), L2 flushed between launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsac_b200 import engine as E
nb = 1024
coords, pix, gt_cv, gt_jp = E.synth_frames(nb)
dc = torch.from_numpy(coords).cuda(); dp = torch.from_numpy(pix).cuda(); dg = torch.from_numpy(gt_jp).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for wd in (1, 0):
    eng = E.Engine(max_frames=nb, write_diffmaps=wd)
    eng.set_stages(E.STAGE_ALL)
    eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
    eng.set_stages(E.STAGE_SCORE)
    ts = []
    for i in range(12):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st); b.record()
        torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    print("write_diffmaps", wd, "variant", os.environ.get("DSAC_K2_VARIANT", "0"), "k_score ms", sum(ts) / len(ts), min(ts))
    del eng
