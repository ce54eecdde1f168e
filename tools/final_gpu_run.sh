#!/bin/bash
# Round-end GPU pass: parity tests of everything touched last, the pageable-caller measurement,
# sanitizer runs over the new kernels, then the rest of the GPU suite with the time that is left.
# Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/final_steps.log; }
stamp start
timeout 120 python -m pytest tests/test_gpu_bank.py tests/test_gpu_convert.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/final_t_changed.log 2>&1; stamp "bank + convert + fuzz tests rc=$?"
timeout 90 python tools/bench_pageable.py --out gpurun_out/bench_pageable.json > gpurun_out/bench_pageable.log 2>&1; stamp "pageable bench rc=$?"
timeout 60 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_bank.py tests/test_gpu_convert.py -m gpu -q \
    -k "repeat_is_read_then_write or batch or pageable_callers" > gpurun_out/sanitizer_memcheck_bank.log 2>&1; stamp "memcheck rc=$?"
timeout 45 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_bank.py -m gpu -q \
    -k "repeat_is_read_then_write" > gpurun_out/sanitizer_racecheck_bank.log 2>&1; stamp "racecheck rc=$?"
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; stamp "smoke rc=$?"
timeout 100 python -m pytest tests -m gpu -x -q --ignore tests/test_gpu_exhaustive.py --ignore tests/test_gpu_bank.py \
    --ignore tests/test_gpu_convert.py --ignore tests/test_gpu_fuzz.py > gpurun_out/final_t_rest.log 2>&1; stamp "rest of the gpu suite rc=$?"
tail -2 gpurun_out/final_t_changed.log gpurun_out/final_t_rest.log
