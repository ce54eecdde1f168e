#!/bin/bash
# Round-end GPU pass: the newest code's parity tests first, then the measurements that are
# still missing, then the whole GPU suite with whatever box time is left.  Everything lands in
# gpurun_out/.
set +e
mkdir -p gpurun_out
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a gpurun_out/final_steps.log; }
stamp start
timeout 240 python -m pytest tests/test_gpu_bank.py -m gpu -x -q > gpurun_out/final_t_bank.log 2>&1; stamp "bank tests rc=$?"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; stamp "smoke rc=$?"
timeout 150 python tools/sweep_bank_repeat.py --out gpurun_out/sweep_bank_repeat.json > gpurun_out/final_sweep_bank_repeat.log 2>&1; stamp "bank repeat sweep rc=$?"
timeout 60 python bench.py --workload bank --fused --steps 200 > gpurun_out/bench_bank_fused_n1.json 2> gpurun_out/bench_bank_fused_n1.err; stamp "bench bank fused rc=$?"
timeout 60 python bench.py --workload bank --fused --graph --steps 200 > gpurun_out/bench_bank_fused_graph_n1.json 2> gpurun_out/bench_bank_fused_graph_n1.err; stamp "bench bank fused graph rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_bank_fused.csv python bench.py --workload bank --fused --steps 5 --warmup 3 > gpurun_out/ncu_bank_fused.log 2>&1; stamp "ncu bank fused rc=$?"
timeout 240 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; stamp "bench n1 rc=$?"
timeout 120 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err; stamp "bench reference rc=$?"
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/final_t_all.log 2>&1; stamp "all gpu tests rc=$?"
tail -3 gpurun_out/final_t_all.log
cat gpurun_out/final_steps.log
