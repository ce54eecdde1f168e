#!/bin/bash
# usage: gpu_final_multi.sh N   -- the N-GPU pass: multi-device tests, the torchrun bench (which also
# carries the one-process leg), the standalone one-process arm, the reference arm.
N=$1
mkdir -p gpurun_out
T=m$N
timeout 900 python -m pytest tests/test_gpu_multi_device.py -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --single-process --gpus $N --steps 50 > gpurun_out/${T}_bench_single.json 2> gpurun_out/${T}_bench_single.err; echo "single exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload sweep --steps 10 > gpurun_out/${T}_bench_sweep.json 2> gpurun_out/${T}_bench_sweep.err
python - <<PY
import json
for f in ('${T}_bench','${T}_bench_single','${T}_bench_ref','${T}_bench_sweep'):
    try:
        b=json.load(open(f'gpurun_out/{f}.json')); e=b.get('e2e') or {}
        print(f, round(b['value'],1), (b.get('roofline') or {}).get('frac'), e.get('value'), e.get('frac_of_link'), e.get('raw_link_gbs_per_rank'), b.get('single_process'))
    except Exception as ex: print(f,'ERR',ex)
PY
nproc; tail -3 gpurun_out/${T}_bench.err
