#!/bin/bash
mkdir -p gpurun_out
T=d5
timeout 1200 python -m pytest tests/test_gpu_bank.py tests/test_gpu_hook.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
timeout 600 python tools/sweep_direct.py --only warp --tag ${T}_sweep_warp > gpurun_out/${T}_sweep.log 2>&1; echo "sweep exit $?"; tail -4 gpurun_out/${T}_sweep.log
timeout 300 python bench.py --workload bank --steps 200 > gpurun_out/${T}_bench_bank_two_calls.json 2> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --steps 200 --external > gpurun_out/${T}_bench_bank_two_calls_external.json 2>> gpurun_out/${T}_bench_bank.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench*.json')):
    try:
        b=json.load(open(f)); r=b.get('roofline') or {}
        print(f, round(b['value'],1), b['ms_per_step'], r.get('frac'))
    except Exception as ex: print(f,'ERR',ex)
PY
