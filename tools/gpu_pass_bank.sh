#!/bin/bash
# Short GPU pass for the batched / bank kernels: parity tests, schedule sweep, bench lines.
set +e
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/bank_steps.log; }
timeout 300 python -m pytest tests/test_gpu_bank.py tests/test_gpu_convert.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/bank_t.log 2>&1; stamp "bank + convert + fuzz tests rc=$?"
timeout 200 python tools/sweep_bank_repeat.py --out gpurun_out/sweep_bank_repeat.json > gpurun_out/sweep_bank_repeat.log 2>&1; stamp "sweep rc=$?"
timeout 60 python bench.py --workload bank --fused --steps 200 > gpurun_out/bench_bank_fused_n1.json 2> gpurun_out/bench_bank_fused_n1.err; stamp "bench fused rc=$?"
timeout 60 python bench.py --workload bank --steps 200 > gpurun_out/bench_bank_n1.json 2> gpurun_out/bench_bank_n1.err; stamp "bench unfused rc=$?"
timeout 60 python bench.py --workload bank --fused --graph --steps 200 > gpurun_out/bench_bank_fused_graph_n1.json 2> gpurun_out/bench_bank_fused_graph_n1.err; stamp "bench fused graph rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_bank_fused.csv python bench.py --workload bank --fused --steps 5 --warmup 3 > gpurun_out/ncu_bank_fused.log 2>&1; stamp "ncu rc=$?"
tail -2 gpurun_out/bank_t.log
