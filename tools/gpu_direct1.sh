#!/bin/bash
# The direct (one tile per CTA) schedules: parity tests over every schedule, then the A/B sweep.
mkdir -p gpurun_out
T=d1
timeout 1200 python -m pytest tests/test_gpu_convert.py tests/test_gpu_round2.py tests/test_gpu_bank.py tests/test_gpu_fuzz.py tests/test_gpu_hook.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -8 gpurun_out/${T}_pytest.log
timeout 900 python tools/sweep_direct.py --tag ${T}_sweep_direct > gpurun_out/${T}_sweep.log 2>&1; echo "sweep exit $?"; tail -40 gpurun_out/${T}_sweep.log
