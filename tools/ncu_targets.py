#!/usr/bin/env python
"""A few launches of each kernel worth profiling, for `ncu --set full` (run under ncu with
--kernel-name / --launch-skip filters; see tools/gpu_final_1gpu.sh).  Arguments: which group."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from sxxcvr_b200 import Bank, Context  # noqa: E402
from sxxcvr_b200.capi import Block  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = Context(0)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
st = side.cuda_stream
n = 1 << 27
src = torch.empty(2 * n, dtype=torch.int32, device="cuda")
cf = torch.empty(2 * n, dtype=torch.float32, device="cuda")
dst = torch.empty(2 * n, dtype=torch.int32, device="cuda")
ctx.synth_frames(src.data_ptr(), 0, n, 1, st)
if which in ("all", "convert"):
    for _ in range(3):
        ctx.convert_rx_buffer(src.data_ptr(), 0, cf.data_ptr(), 0, n, st)
        ctx.convert_tx_buffer(cf.data_ptr(), 0, dst.data_ptr(), 0, n, 1e-6, st)
if which in ("all", "batch"):
    nb, m = 1024, n // 1024
    arr = (Block * nb)(*[Block(src.data_ptr() + 8 * m * b, cf.data_ptr() + 8 * m * b, m, 0.0, 0) for b in range(nb)])
    d_list = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).cuda()
    for _ in range(3):
        ctx.convert_batch("rx", d_list.data_ptr(), on_device=True, max_length=m, stream=st, nblocks=nb)
if which in ("all", "loopback"):
    for _ in range(3):
        ctx.convert_loopback(src.data_ptr(), cf.data_ptr(), dst.data_ptr(), n, 1e-6, st)
if which in ("all", "bank"):
    S, P = 65536, 256
    for variant in (0, 100):
        ctx.set_option("bank_repeat_variant", variant)
        with Bank(ctx, S, P, 75000.0, 0.0, 7) as bank:
            for _ in range(3):
                bank.repeat(cf.data_ptr(), 10_240_000, st)
            torch.cuda.synchronize()
torch.cuda.synchronize()
ctx.close()
