"""Single-frame latency (BASELINE config 2: 1 frame x 256 hypotheses, sample + score + soft-argmax), CUDA events over 200 calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsac_b200 import engine as E
for T in (1, 8):
    coords, pix, gt_cv, gt_jp = E.synth_frames(1, n_streams=T)
    eng = E.Engine(max_frames=1, n_streams=T, write_diffmaps=0)
    eng.set_stages(E.STAGE_SAMPLE | E.STAGE_SCORE)
    dc = torch.from_numpy(coords).cuda(); dp = torch.from_numpy(pix).cuda(); dg = torch.from_numpy(gt_jp).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(10): eng.forward_device(1, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): eng.forward_device(1, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
    b.record(); torch.cuda.synchronize()
    print("T=%d latency %.1f us" % (T, a.elapsed_time(b) * 1e3 / 200), flush=True)
    eng.close()
