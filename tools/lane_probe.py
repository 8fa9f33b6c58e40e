"""Development aid: would splitting ONE 1024-frame step into P concurrent sub-batches (P engines on P streams, device-resident
inputs, step-isolated with an L2 flush like bench.py's `value`) beat the single pass?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dsac_b200 import engine as E
NF = 1024
coords, pix, gt_cv, gt_jp = E.synth_frames(NF)
d_c, d_p, d_g = torch.from_numpy(coords).cuda(), torch.from_numpy(pix).cuda(), torch.from_numpy(gt_jp).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for P in (1, 2, 3, 4, 6):
    bounds = [NF * i // P for i in range(P + 1)]
    engs = [E.Engine(max_frames=bounds[i + 1] - bounds[i]) for i in range(P)]
    streams = [torch.cuda.Stream() for _ in range(P)]
    def step():
        ev = torch.cuda.Event(); ev.record()
        for i in range(P):
            lo, n = bounds[i], bounds[i + 1] - bounds[i]
            streams[i].wait_event(ev)
            engs[i].forward_device(n, d_c[lo:].data_ptr(), d_p[lo:].data_ptr(), 0, d_g[lo:].data_ptr(), lo, streams[i].cuda_stream)
        for s in streams: torch.cuda.current_stream().wait_stream(s)
    for _ in range(4): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    print("P=%d sub-batches: %.3f ms per 1024-frame step (min %.3f)" % (P, float(np.mean(ts)), min(ts)), flush=True)
    for e in engs: e.close()
