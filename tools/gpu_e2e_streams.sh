#!/bin/bash
mkdir -p gpurun_out
for lg in 27 25 23; do
  timeout 300 python bench.py --steps 20 --no-rows --no-cpu-baseline --min-seconds 0 --e2e-log2-frames $lg > gpurun_out/e2e_lg_$lg.json 2> gpurun_out/e2e_lg_$lg.err
  python - <<PY
import json
b=json.load(open('gpurun_out/e2e_lg_$lg.json')); e=b['e2e']
print($lg, round(e['value'],1), e['frac_of_link'], e['raw_link_gbs_per_rank']['both_each_way_gbs'], e['frames_per_call'])
PY
done
