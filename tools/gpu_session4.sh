#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_exhaustive.py > gpurun_out/s4_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s4_pytest.log
tail -12 gpurun_out/s4_pytest.log
python tools/probe_batch.py 6 > gpurun_out/s4_probe_batch.log 2>&1; cat gpurun_out/s4_probe_batch.log
timeout 600 python tools/sweep_round2.py --only batched,loopback --tag s4_sweep > gpurun_out/s4_sweep.log 2>&1; tail -8 gpurun_out/s4_sweep.log
timeout 600 python bench.py > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/s4_bench_ref.json 2>> gpurun_out/s4_bench.err
timeout 300 python bench.py --workload sweep --steps 10 > gpurun_out/s4_bench_sweep.json 2> gpurun_out/s4_bench_sweep.err; echo "sweep exit $?"
timeout 600 python bench.py --workload group > gpurun_out/s4_bench_group.json 2> gpurun_out/s4_bench_group.err; echo "group exit $?"
timeout 300 python bench.py --single-process --gpus 1 --steps 50 > gpurun_out/s4_bench_single.json 2> gpurun_out/s4_bench_single.err; echo "single exit $?"
timeout 300 python bench.py --workload bank --fused --graph --steps 200 > gpurun_out/s4_bench_bank.json 2> gpurun_out/s4_bench_bank.err; echo "bank exit $?"
python - <<'PY'
import json
for f in ('s4_bench','s4_bench_ref','s4_bench_sweep','s4_bench_group','s4_bench_single','s4_bench_bank'):
    try:
        b=json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(b['value'],1), (b.get('roofline') or {}).get('frac'), (b.get('e2e') or {}).get('value'))
    except Exception as ex: print(f,'ERR',ex)
PY
# ncu: launch list of the bench command, then full captures of the hot kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 2 --warmup 1 --no-rows --no-cpu-baseline --min-seconds 0 > gpurun_out/s4_launches.log 2>&1
for g in convert batch loopback bank; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bulk_convert_kernel|bulk_batch_kernel|bulk_loopback_kernel|bank_repeat' --launch-skip 2 --launch-count 4 -f -o gpurun_out/s4_ncu_$g python tools/ncu_targets.py $g > gpurun_out/s4_ncu_$g.log 2>&1
  echo "ncu $g exit $?"
done
ls -la gpurun_out/*.ncu-rep
