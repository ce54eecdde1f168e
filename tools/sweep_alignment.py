#!/usr/bin/env python
"""Throughput of a 2^27-frame conversion as a function of the frame offsets of source and destination."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402

ctx = Context(0)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
st = side.cuda_stream
n = 1 << 27
i2s = torch.empty(2 * n + 64, dtype=torch.int32, device="cuda")
cf = torch.empty(2 * n + 64, dtype=torch.float32, device="cuda")
out = torch.empty(2 * n + 64, dtype=torch.int32, device="cuda")
ctx.synth_frames(i2s.data_ptr(), 0, n + 32, 1, st)
ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n + 32, st)
if True:
    for so, do, extra in ((0, 0, 0), (1, 1, 0), (2, 2, 0), (3, 1, 0), (1, 0, 0), (0, 1, 0), (3, 2, 0), (0, 0, 4)):
        row = []
        for name, fn in (("rx", lambda: ctx.convert_rx_buffer(i2s.data_ptr() + extra, so, cf.data_ptr() + extra, do, n, st)),
                         ("tx", lambda: ctx.convert_tx_buffer(cf.data_ptr() + extra, so, out.data_ptr() + extra, do, n, 1e-6, st))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn()
            b.record()
            torch.cuda.synchronize()
            row.append(16 * n / (a.elapsed_time(b) / 10) / 1e6)
        print(f"src_offset={so} dest_offset={do} byte_skew={extra}: RX {row[0]:6.0f} GB/s  TX {row[1]:6.0f} GB/s", flush=True)
