#!/bin/bash
# Last check of the final binary on one B200: what the driver runs (GPU suite, smoke(), both bench
# arms) plus the bank arms and memcheck over the large-bank schedules.
mkdir -p gpurun_out
T=${1:-f3}
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err ) 2>&1 | grep real
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_steps20.json 2> gpurun_out/${T}_bench.err ) 2>&1 | grep real
timeout 300 python bench.py --workload bank --fused --graph --steps 200 > gpurun_out/${T}_bench_bank_fused_graph.json 2> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --steps 200 > gpurun_out/${T}_bench_bank_two_calls.json 2>> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --steps 200 --external > gpurun_out/${T}_bench_bank_two_calls_external.json 2>> gpurun_out/${T}_bench_bank.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench*.json')):
    try:
        b=json.load(open(f)); e=b.get('e2e') or {}; r=b.get('roofline') or {}
        print(f, round(b['value'],1), r.get('frac'), r.get('frac_of_write_only_ceiling'), e.get('value'), e.get('frac_of_link'))
    except Exception as ex: print(f,'ERR',ex)
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest "tests/test_gpu_bank.py::test_large_bank_default_schedules_and_grid_options" -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_sanitizer_memcheck_bank_split.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/${T}_sanitizer_memcheck_bank_split.log
