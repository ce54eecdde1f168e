#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_hook.py tests/test_gpu_group.py tests/test_gpu_bank.py tests/test_gpu_convert.py tests/test_gpu_stream.py tests/test_gpu_stream_fuzz.py tests/test_gpu_compat.py tests/test_gpu_multi_device.py tests/test_gpu_fuzz.py tests/test_bench_contract.py -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/s3_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s3_pytest.log
tail -30 gpurun_out/s3_pytest.log
timeout 1200 python tools/sweep_round2.py --tag s3_sweep > gpurun_out/s3_sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/s3_sweep.log
tail -3 gpurun_out/s3_sweep.log
for kind in pageable pinned pin; do
  timeout 600 python bench.py --no-rows --no-cpu-baseline --steps 10 --e2e-buffers $kind > gpurun_out/s3_bench_$kind.json 2> gpurun_out/s3_bench_$kind.err
  echo "bench $kind exit $?"
done
for streams in 1 2 4; do
  timeout 600 python bench.py --no-rows --no-cpu-baseline --steps 10 --e2e-buffers pinned --e2e-streams $streams > gpurun_out/s3_bench_pinned_s$streams.json 2>> gpurun_out/s3_bench_pinned.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s3_bench_*.json')):
    try:
        b=json.load(open(f)); e=b['e2e']
        print(f, round(e['value'],1), e['streams_per_gpu'], e['frac_of_link'], e['raw_link_gbs_per_rank']['both_each_way_gbs'])
    except Exception as ex: print(f, 'ERR', ex)
PY
