#!/usr/bin/env python
"""Round-robin vs contiguous tile assignment of the bulk kernel, sustained back-to-back launches."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402

ctx = Context(0)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
st = side.cuda_stream
res = []
for lg in (27, 24):
    n = 1 << lg
    i2s = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    cf = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    out = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    ctx.synth_frames(i2s.data_ptr(), 0, n, 1, st)
    ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)
    for rep in range(3):
        for contiguous in (0, 1):
            ctx.set_option("bulk_contiguous", contiguous)
            for _ in range(3):
                ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)
                ctx.convert_tx_buffer(cf.data_ptr(), 0, out.data_ptr(), 0, n, 1e-6, st)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(30):
                ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)
                ctx.convert_tx_buffer(cf.data_ptr(), 0, out.data_ptr(), 0, n, 1e-6, st)
            b.record()
            torch.cuda.synchronize()
            gbs = 2 * 16 * n / (a.elapsed_time(b) / 30) / 1e6
            res.append(dict(log2_frames=lg, contiguous=contiguous, rep=rep, gbs=gbs))
            print(f"2^{lg} contiguous={contiguous} rep{rep}: {gbs:7.0f} GB/s (RX+TX pairs, back to back)", flush=True)
    del i2s, cf, out
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sweep_assignment.json").write_text(json.dumps(res, indent=1))
