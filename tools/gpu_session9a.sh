#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bank.py tests/test_gpu_hook.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/s9a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s9a_pytest.log; tail -15 gpurun_out/s9a_pytest.log
timeout 900 python tools/sweep_round2.py --only bank,ext --tag s9a_sweep_bank > gpurun_out/s9a_sweep.log 2>&1; cat gpurun_out/s9a_sweep.log | cut -c1-1000
for v in 100 400; do
timeout 300 python bench.py --workload bank --fused --graph --steps 200 --repeat-variant $v > gpurun_out/s9a_bench_bank_v$v.json 2> gpurun_out/s9a_bench_bank.err
timeout 300 python bench.py --workload bank --fused --graph --external --steps 200 --repeat-variant $v > gpurun_out/s9a_bench_bank_ext_v$v.json 2>> gpurun_out/s9a_bench_bank.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s9a_bench_*.json')):
    try:
        b=json.load(open(f)); print(f, round(b['value'],1), b['ms_per_step'], (b.get('roofline') or {}).get('frac'))
    except Exception as ex: print(f,'ERR',ex)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bank_repeat_bulk' --launch-skip 2 --launch-count 2 -f -o gpurun_out/s9a_ncu_bank_bulk python bench.py --workload bank --fused --steps 5 --repeat-variant 400 > gpurun_out/s9a_ncu.log 2>&1; echo "ncu exit $?"
