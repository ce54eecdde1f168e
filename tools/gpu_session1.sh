#!/bin/bash
# Round-2 GPU session 1: correctness of everything new, then the sweeps, then bench.py.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/s1_gpu.txt 2>&1
nproc >> gpurun_out/s1_gpu.txt; lscpu | grep -i "model name\|socket\|numa" >> gpurun_out/s1_gpu.txt
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_bank.py tests/test_gpu_convert.py tests/test_gpu_stream.py tests/test_gpu_stream_fuzz.py tests/test_gpu_compat.py tests/test_gpu_fuzz.py tests/test_bench_contract.py -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/s1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s1_pytest.log
tail -40 gpurun_out/s1_pytest.log
timeout 900 python tools/sweep_round2.py > gpurun_out/s1_sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/s1_sweep.log
tail -5 gpurun_out/s1_sweep.log
timeout 600 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/s1_bench.json
tail -5 gpurun_out/s1_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/s1_bench_ref.json 2>> gpurun_out/s1_bench.err
tail -c 600 gpurun_out/s1_bench_ref.json
