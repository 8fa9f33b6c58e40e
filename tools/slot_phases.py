"""Development aid: thread-0 cycle split of the generator (k1_slot) for one frame / a full batch (DSAC_K1_TIMERS=1)."""
import os, sys
os.environ["DSAC_K1_TIMERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsac_b200 import engine as E
nb = int(os.environ.get("NB", "1"))
coords, pix, gt_cv, gt_jp = E.synth_frames(nb)
eng = E.Engine(max_frames=nb, write_diffmaps=0)
eng.set_stages(E.STAGE_SAMPLE)
dc = torch.from_numpy(coords).cuda(); dp = torch.from_numpy(pix).cuda(); dg = torch.from_numpy(gt_jp).cuda()
st = torch.cuda.current_stream().cuda_stream
for _ in range(10): eng.forward_device(nb, dc.data_ptr(), dp.data_ptr(), 0, dg.data_ptr(), 0, st)
torch.cuda.synchronize()
print("10 passes over", nb, "frame(s)", flush=True)
eng.close()
