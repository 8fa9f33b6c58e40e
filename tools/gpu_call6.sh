# call 6: k_sample A2 (pair parser, shuffle scan), backward kernels over the whole GPU; tests; timers; launch list of the training rounds
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=180 -rf > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c6_pytest.log
rm -f gpurun_out/sweep.jsonl
SWEEP_STEPS=20 timeout 600 python tools/sweep.py run > gpurun_out/c6_sweep.log 2>&1; echo "sweep rc=$?"
DSAC_K1_TIMERS=1 NB=1024 REPS=3 timeout 120 python tools/prof_driver.py 2>&1 | grep -i "cycles" > gpurun_out/c6_k1_timers.txt; cat gpurun_out/c6_k1_timers.txt
timeout 120 python tools/dsac_probe.py > gpurun_out/c6_dsac_probe.txt 2>&1; cat gpurun_out/c6_dsac_probe.txt
REPS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/c6_dsac_launches.csv python tools/dsac_probe.py > /dev/null 2>&1; python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/c6_dsac_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size")
for r in rows[1:]:
    print(r[ki][:40], r[gi], r[vi])
PY
