# sweep #5: k_sample with one barrier per 624-word regeneration, smaller super-rounds / larger L1; final k_score; pooled backward buffers
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=180 -rf > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c5_pytest.log
rm -f gpurun_out/sweep.jsonl
SWEEP_STEPS=20 timeout 600 python tools/sweep.py run > gpurun_out/c5_sweep.log 2>&1; echo "sweep rc=$?"
DSAC_K1_TIMERS=1 NB=1024 REPS=3 timeout 120 python tools/prof_driver.py 2>&1 | grep -i "cycles" > gpurun_out/c5_k1_timers.txt; cat gpurun_out/c5_k1_timers.txt
