"""Development aid: time the step / the sampler of one library build under a set of run-time knobs (environment), one
subprocess per configuration; prints one JSON line each.   python tools/knob_probe.py run  (on the GPU box)
Configurations: tools/knob_probe.json = [{"name": .., "lib": "variants/x.so" or null, "env": {..}}, ...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np, torch
    import dsac_b200.engine as E
    NF = int(os.environ.get("SWEEP_FRAMES", "1024")); STEPS = int(os.environ.get("SWEEP_STEPS", "10"))
    coords, pix, gt_cv, gt_jp = E.synth_frames(NF)
    eng = E.Engine(max_frames=NF)
    stream = torch.cuda.current_stream().cuda_stream
    d_c, d_p, d_g = torch.from_numpy(coords).cuda(), torch.from_numpy(pix).cuda(), torch.from_numpy(gt_jp).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def step():
        eng.forward_device(NF, d_c.data_ptr(), d_p.data_ptr(), 0, d_g.data_ptr(), 0, stream)
    def timed(k):
        tot = []
        for _ in range(k):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); b.synchronize()
            tot.append(a.elapsed_time(b))
        return round(float(np.mean(tot)), 4), round(float(np.min(tot)), 4)
    out = {"name": os.environ.get("KNOB_NAME", "?")}
    for _ in range(4): step()
    torch.cuda.synchronize()
    out["step_ms"] = timed(STEPS)
    r = eng.fetch(NF)
    out["chk"] = [int(r.img_idx.astype(np.int64).sum()), int(r.cand_idx.astype(np.int64).sum()), int(r.inlier_map.sum()), float(np.abs(r.ref_pose).sum())]
    eng.set_stages(E.STAGE_SAMPLE); step(); step(); torch.cuda.synchronize()
    out["sampler_ms"] = timed(STEPS)
    out["launches_per_pass"] = None
    eng.close()
    print("KNOB " + json.dumps(out))


def run():
    cfgs = json.load(open(os.path.join(ROOT, "tools", "knob_probe.json")))
    outp = os.path.join(ROOT, "gpurun_out", "knobs.jsonl"); os.makedirs(os.path.dirname(outp), exist_ok=True)
    with open(outp, "a") as fh:
        for c in cfgs:
            env = dict(os.environ, KNOB_NAME=c["name"], DSAC_SKIP_BUILD="1", **{k: str(v) for k, v in c.get("env", {}).items()})
            if c.get("lib"): env["DSAC_B200_LIB"] = os.path.join(ROOT, c["lib"])
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True, timeout=240)
                lines = [x for x in r.stdout.splitlines() if x.startswith("KNOB ")]
                msg = lines[-1][5:] if lines else json.dumps({"name": c["name"], "error": (r.stderr or r.stdout)[-500:]})
            except subprocess.TimeoutExpired:
                msg = json.dumps({"name": c["name"], "error": "timeout"})
            fh.write(msg + "\n"); fh.flush(); print(msg)


if __name__ == "__main__":
    {"one": one, "run": run}[sys.argv[1]]()
