#!/bin/bash
# Round-2 final pass on one B200: the whole GPU suite (exhaustive sweeps included), smoke(), every
# bench arm, the A/B sweeps, the ncu launch list and full captures, and the sanitizers over what is new.
mkdir -p gpurun_out
T=${1:-f2}
QUICK=${2:-}   # "quick": benches, direct sweep and ncu only (the suite and the sanitizers ran in tools/gpu_final_check.sh)
if [ -z "$QUICK" ]; then
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_steps20.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --workload sweep --steps 20 > gpurun_out/${T}_bench_sweep.json 2> gpurun_out/${T}_bench_sweep.err
timeout 600 python bench.py --workload group > gpurun_out/${T}_bench_group.json 2> gpurun_out/${T}_bench_group.err
timeout 300 python bench.py --single-process --gpus 1 --steps 50 > gpurun_out/${T}_bench_single.json 2> gpurun_out/${T}_bench_single.err
timeout 300 python bench.py --workload bank --fused --graph --steps 200 > gpurun_out/${T}_bench_bank_fused_graph.json 2> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --fused --steps 200 > gpurun_out/${T}_bench_bank_fused.json 2>> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --fused --graph --external --steps 200 > gpurun_out/${T}_bench_bank_fused_graph_external.json 2>> gpurun_out/${T}_bench_bank.err
timeout 300 python bench.py --workload bank --steps 200 > gpurun_out/${T}_bench_bank_two_calls.json 2>> gpurun_out/${T}_bench_bank.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench*.json')):
    try:
        b=json.load(open(f)); e=b.get('e2e') or {}; r=b.get('roofline') or {}
        print(f, round(b['value'],1), r.get('frac'), r.get('frac_of_write_only_ceiling'), e.get('value'), e.get('frac_of_link'))
    except Exception as ex: print(f,'ERR',ex)
PY
timeout 900 python tools/sweep_direct.py --tag ${T}_sweep_direct > gpurun_out/${T}_sweep_direct.log 2>&1
[ -z "$QUICK" ] && timeout 900 python tools/sweep_round2.py --only host,plugin --tag ${T}_sweep_host_plugin > gpurun_out/${T}_sweep_host_plugin.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-rows --no-cpu-baseline --min-seconds 0 > gpurun_out/${T}_launches.log 2>&1
for g in convert batch loopback bank; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stream_convert_kernel|batch_direct_kernel|loopback_kernel|bank_repeat|bank_plan_repeat' --launch-skip 2 --launch-count 4 -f -o gpurun_out/${T}_ncu_$g python tools/ncu_targets.py $g > gpurun_out/${T}_ncu_$g.log 2>&1
done
ls gpurun_out/${T}_ncu_*.ncu-rep
[ -n "$QUICK" ] && exit 0
SAN="tests/test_gpu_round2.py::test_batched_mid_size_blocks tests/test_gpu_round2.py::test_loopback_schedules tests/test_gpu_convert.py::test_frame_offsets_and_misaligned_views tests/test_gpu_bank.py::test_repeat_with_blocks_that_straddle_ring_slices tests/test_gpu_bank.py::test_ingested_frames_replace_the_synthetic_capture tests/test_gpu_bank.py::test_repeat_is_read_then_write tests/test_gpu_hook.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $SAN -m gpu -q -x -p no:cacheprovider -k "not 32771" > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/${T}_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_round2.py::test_batched_mid_size_blocks tests/test_gpu_round2.py::test_loopback_schedules "tests/test_gpu_bank.py::test_repeat_with_blocks_that_straddle_ring_slices" -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/${T}_sanitizer_racecheck.log
