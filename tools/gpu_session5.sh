#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_bank.py tests/test_gpu_stream.py tests/test_gpu_compat.py tests/test_gpu_hook.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/s5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s5_pytest.log; tail -5 gpurun_out/s5_pytest.log; grep "group of" gpurun_out/s5_pytest.log
timeout 900 python tools/sweep_round2.py --only bank --tag s5_sweep_bank > gpurun_out/s5_sweep.log 2>&1; cat gpurun_out/s5_sweep.log | cut -c1-1200
timeout 600 python bench.py --workload group > gpurun_out/s5_bench_group.json 2> gpurun_out/s5_bench_group.err; echo "group exit $?"
for v in 0 100 204 300; do
timeout 300 python bench.py --workload bank --fused --graph --external --steps 200 --repeat-variant $v > gpurun_out/s5_bench_bank_ext_v$v.json 2> gpurun_out/s5_bench_bank_ext.err; done
for st in 8 12 16; do
timeout 300 python bench.py --no-rows --no-cpu-baseline --min-seconds 0 --steps 10 --e2e-streams $st > gpurun_out/s5_bench_st$st.json 2> gpurun_out/s5_bench_st.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s5_bench_*.json')):
    try:
        b=json.load(open(f)); e=b.get('e2e') or {}
        print(f, round(b['value'],1), (b.get('roofline') or {}).get('frac'), e.get('value'), e.get('frac_of_link'))
        if 'group' in b:
            for r in b['group']: print('   ', r)
    except Exception as ex: print(f,'ERR',ex)
PY
