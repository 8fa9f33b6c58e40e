# one gpurun call: smoke -> variant sweep -> GPU tests -> sanitizer on a small batch
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -1; nproc
timeout 300 python __graft_entry__.py --smoke > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c1_smoke.log
rm -f gpurun_out/sweep.jsonl
timeout 900 python tools/sweep.py run > gpurun_out/c1_sweep.log 2>&1; echo "sweep rc=$?"
timeout 700 python -m pytest tests -m gpu -q --timeout=180 -rf > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c1_pytest.log
for tool in racecheck memcheck; do NB=3 REPS=1 timeout 200 compute-sanitizer --tool $tool python tools/prof_driver.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error" | head -5 | sed "s/^/$tool: /"; done > gpurun_out/c1_sanitizer.txt 2>&1
cat gpurun_out/c1_sanitizer.txt
