# sweep #4: k_score with 16-byte stores (pairs / one-reciprocal sigmoid), carve-out attributes, DSAC round probe, 3 engines in flight
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=180 -rf > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c4_pytest.log
timeout 120 python tools/dsac_probe.py > gpurun_out/c4_dsac_probe.txt 2>&1; cat gpurun_out/c4_dsac_probe.txt
BIG=1 timeout 120 python tools/dsac_probe.py > gpurun_out/c4_dsac_probe_big.txt 2>&1; cat gpurun_out/c4_dsac_probe_big.txt
REPS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/c4_dsac_launches.csv python tools/dsac_probe.py > /dev/null 2>&1; python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/c4_dsac_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
for r in rows[1:]:
    print(r[ki][:40], r[gi] if gi is not None else "", r[vi])
PY
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py run > gpurun_out/c4_sweep.log 2>&1; echo "sweep rc=$?"
