cd $GRAFT_REPO_ROOT; export DSAC_SKIP_BUILD=1; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -x --timeout=300 -k "speculative" 2>&1 | tail -2
timeout 300 python tools/sampler_breakdown.py 2>&1 | grep "^n=" | head -3
DSAC_K1_SPEC=0 timeout 300 python tools/sampler_breakdown.py 2>&1 | grep "^n=" | head -3 | sed 's/^/spec=0: /'
for n in 8 16; do SWEEP_FRAMES=$n SWEEP_STEPS=20 timeout 300 python tools/knob_probe.py one 2>&1 | tail -1; DSAC_K1_SPEC=0 SWEEP_FRAMES=$n SWEEP_STEPS=20 timeout 300 python tools/knob_probe.py one 2>&1 | tail -1 | sed 's/^/spec=0: /'; done
