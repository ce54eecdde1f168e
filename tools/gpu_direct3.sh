#!/bin/bash
# Direct schedules as defaults: the whole GPU suite, the bench arms they change, the sweep.
mkdir -p gpurun_out
T=d3
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
timeout 300 python bench.py --workload sweep --steps 20 > gpurun_out/${T}_bench_sweep.json 2> gpurun_out/${T}_bench_sweep.err
timeout 300 python bench.py --workload bank --fused --graph --steps 200 > gpurun_out/${T}_bench_bank_fused_graph.json 2> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --fused --steps 200 > gpurun_out/${T}_bench_bank_fused.json 2>> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --fused --graph --external --steps 200 > gpurun_out/${T}_bench_bank_fused_graph_external.json 2>> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --single-process --gpus 1 --steps 50 > gpurun_out/${T}_bench_single.json 2> gpurun_out/${T}_bench_single.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/d3_bench*.json')):
    try:
        b=json.load(open(f)); e=b.get('e2e') or {}
        print(f, round(b['value'],1), (b.get('roofline') or {}).get('frac'), (b.get('roofline') or {}).get('frac_of_write_only_ceiling'), e.get('value'), e.get('frac_of_link'))
    except Exception as ex: print(f,'ERR',ex)
PY
timeout 900 python tools/sweep_direct.py --only convert,bank --tag ${T}_sweep_direct > gpurun_out/${T}_sweep.log 2>&1; echo "sweep exit $?"; tail -12 gpurun_out/${T}_sweep.log
