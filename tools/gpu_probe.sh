#!/bin/bash
mkdir -p gpurun_out
python tools/probe_batch.py 8 > gpurun_out/probe_batch.log 2>&1
cat gpurun_out/probe_batch.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/probe_batch_ncu.csv python tools/probe_batch.py 2 > gpurun_out/probe_batch_ncu.log 2>&1
grep -c bulk_batch gpurun_out/probe_batch_ncu.csv
