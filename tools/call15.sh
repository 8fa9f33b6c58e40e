cd $GRAFT_REPO_ROOT; export DSAC_SKIP_BUILD=1; mkdir -p gpurun_out
DSAC_K1_DEBUG=1 timeout 120 python tools/latency_probe.py 2>&1 | grep -E "latency|k1_spec" | head -4
DSAC_K1_SOLVE4=0 timeout 120 python tools/latency_probe.py 2>&1 | grep -E "latency" | sed 's/^/solve4=0: /'
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_configs.py tests/test_gpu_backward.py -m gpu -q -x --timeout=300 2>&1 | tail -3
timeout 300 python tools/sampler_breakdown.py 2>&1 | grep "^n=" | head -3
NB=1 REPS=3 WRITE_DM=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c16_n1.csv python tools/prof_driver.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c16_n1.csv')) if len(r)>10 and r[0].isdigit()]
n=len(rows)//3
for r in rows[-n:]: print("%-40s %8.1f us"%(r[4].split('(')[0][:40], float(r[-1])/1e3))
PY
SWEEP_STEPS=20 timeout 300 python tools/knob_probe.py one 2>&1 | tail -1
SWEEP_FRAMES=128 SWEEP_STEPS=20 timeout 300 python tools/knob_probe.py one 2>&1 | tail -1
