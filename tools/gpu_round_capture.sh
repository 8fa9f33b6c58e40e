set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r01b_pytest.log 2>&1; tail -2 gpurun_out/r01b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01b_bench_line.json 2> gpurun_out/r01b_bench.err; tail -c 600 gpurun_out/r01b_bench_line.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r01b_ncu_bench.log 2>&1
for k in k_score k_sample k_refine; do timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -s 2 -f -o gpurun_out/${k}_r01b python tools/prof_driver.py > gpurun_out/ncu_$k.log 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gather -c 1 -s 2 -f -o gpurun_out/k_gather_r01b python tools/gather_probe.py > gpurun_out/ncu_k_gather.log 2>&1
for tool in memcheck racecheck initcheck; do NB=5 REPS=1 timeout 500 compute-sanitizer --tool $tool python tools/prof_driver.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY" | sed "s/^/$tool fwd: /"; done > gpurun_out/sanitizer_r01b.txt 2>&1
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_backward.py -m gpu -q -k "dsac_variant_backward_matches_oracle and 16" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | sed "s/^/memcheck backward_dsac: /" >> gpurun_out/sanitizer_r01b.txt
cat gpurun_out/sanitizer_r01b.txt
ls -la gpurun_out | tail -12
