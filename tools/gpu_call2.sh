# sweep #3: fixed k_score, k_refine with staged frame data + squared-domain inlier test; GPU tests; ncu captures
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=180 -rf > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c3_pytest.log
rm -f gpurun_out/sweep.jsonl
timeout 600 python tools/sweep.py run > gpurun_out/c3_sweep.log 2>&1; echo "sweep rc=$?"

for k in k_score k_refine; do DSAC_TAIL_SPLIT=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -s 2 -f -o gpurun_out/${k}_c3 python tools/prof_driver.py > gpurun_out/c3_ncu_$k.log 2>&1; echo "ncu $k rc=$?"; done
ls -la gpurun_out | tail -12
