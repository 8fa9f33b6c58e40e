"""Development aid: per-kernel event times of the split sampler (dsac_sampler_profile: one stream, no overlap) for a few
batch sizes / stream counts.   python tools/sampler_breakdown.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsac_b200 import engine as E
for nb, T in ((1, 1), (1, 8), (16, 1), (128, 1), (1024, 1)):
    coords, pix, gt_cv, gt_jp = E.synth_frames(nb, n_streams=T)
    eng = E.Engine(max_frames=nb, n_streams=T, write_diffmaps=0)
    eng.set_stages(E.STAGE_SAMPLE)
    dc = torch.from_numpy(coords).cuda(); dp = torch.from_numpy(pix).cuda(); dg = torch.from_numpy(gt_jp).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5): eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
    b.record(); torch.cuda.synchronize()
    tot = a.elapsed_time(b) / 20
    eng.sampler_profile(True)
    acc = [0.0] * 4
    for _ in range(5):
        eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st); torch.cuda.synchronize()
        ms, cnt = eng.sampler_profile_read()
        acc = [x + y / 5 for x, y in zip(acc, ms)]
    print("n=%d T=%d  sampler %.1f us (back to back, overlapped);  profiled: gen+select %.1f  filter %.1f  solve %.1f  tail %.1f us;  candidates %d flagged %d accepted %d rounds %d  launches/pass %d"
          % (nb, T, tot * 1e3, acc[0] * 1e3, acc[1] * 1e3, acc[2] * 1e3, acc[3] * 1e3, cnt[0], cnt[1], cnt[2], cnt[3], eng.launches // 31), flush=True)
    eng.close()
