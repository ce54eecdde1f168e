#!/usr/bin/env python
"""Per-launch look at the batched kernels: every shape a few times, one CUDA event pair per call
(prints every repetition, so outliers show), meant to be run plain and under
`ncu --metrics gpu__time_duration.sum` for the kernels' own durations."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from sxxcvr_b200 import Context  # noqa: E402
from sxxcvr_b200.capi import Block  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = Context(0)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
st = side.cuda_stream
total = 1 << 27
src = torch.empty(2 * total, dtype=torch.int32, device="cuda")
cf = torch.empty(2 * total, dtype=torch.float32, device="cuda")
ctx.synth_frames(src.data_ptr(), 0, total, 1, st)
for log2n in (17, 19, 21, 23, 25):
    n = 1 << log2n
    nb = total // n
    arr = (Block * nb)(*[Block(src.data_ptr() + 8 * n * b, cf.data_ptr() + 8 * n * b, n, 0.0, 0) for b in range(nb)])
    d_list = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).cuda()
    for variant in (0, 1):
        ctx.set_option("batch_variant", variant)
        times = []
        for r in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(side)
            ctx.convert_batch("rx", d_list.data_ptr(), on_device=True, max_length=n, stream=st, nblocks=nb)
            b.record(side)
            torch.cuda.synchronize()
            times.append(round(a.elapsed_time(b) * 1e3, 1))
        print(json.dumps({"blocks": nb, "variant": "bulk" if variant == 0 else "slices", "us_per_call": times,
                          "best_gbs": round(16 * total / (min(times) * 1e-6) / 1e9, 1)}), flush=True)
ctx.close()
