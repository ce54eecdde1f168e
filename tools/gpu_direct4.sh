#!/bin/bash
# warp-per-stream grids; ncu launch list and full captures of the direct kernels
mkdir -p gpurun_out
T=d4
timeout 600 python tools/sweep_direct.py --only warp --tag ${T}_sweep_warp > gpurun_out/${T}_sweep.log 2>&1; echo "sweep exit $?"; tail -4 gpurun_out/${T}_sweep.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-rows --no-cpu-baseline --min-seconds 0 > gpurun_out/${T}_launches.log 2>&1
for g in convert batch loopback bank; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stream_convert_kernel|batch_direct_kernel|loopback_kernel|bank_repeat|bank_plan_repeat' --launch-skip 2 --launch-count 4 -f -o gpurun_out/${T}_ncu_$g python tools/ncu_targets.py $g > gpurun_out/${T}_ncu_$g.log 2>&1
done
ls -la gpurun_out/${T}_ncu_*.ncu-rep
