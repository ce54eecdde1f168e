#!/usr/bin/env python
"""Which schedule wins for small and medium blocks (2^13 .. 2^24 frames)?  Back-to-back launches
(50 per measurement) between two CUDA events, so per-launch cost includes what a stream of such
calls actually pays."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402

CONFIGS = [
    ("vec128 u4 b256 auto", dict(v=1, unroll=4, block=256)),
    ("vec256 u2 b256 auto", dict(v=2, unroll=2, block=256)),
    ("vec256 u4 b256 auto", dict(v=2, unroll=4, block=256)),
    ("vec256 u8 b512 c8", dict(v=2, unroll=8, block=512, ctas_per_sm=8)),
    ("bulk 2048x4 b256", dict(v=3, bulk_tile=2048, bulk_stages=4, block=256)),
    ("bulk 1024x4 b256", dict(v=3, bulk_tile=1024, bulk_stages=4, block=256)),
    ("bulk 512x4 b256 c2", dict(v=3, bulk_tile=512, bulk_stages=4, block=256, ctas_per_sm=2)),
    ("bulk 512x4 b128 c4", dict(v=3, bulk_tile=512, bulk_stages=4, block=128, ctas_per_sm=4)),
]


def main():
    ctx = Context(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    st = side.cuda_stream
    nmax = 1 << 24
    i2s = torch.empty(2 * nmax, dtype=torch.int32, device="cuda")
    cf = torch.empty(2 * nmax, dtype=torch.float32, device="cuda")
    out_i = torch.empty(2 * nmax, dtype=torch.int32, device="cuda")
    ctx.synth_frames(i2s.data_ptr(), 0, nmax, 1, st)
    ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, nmax, st)
    res = []
    for direction in ("rx", "tx"):
        print(direction, " " * 20, " ".join(f"2^{lg:<6d}" for lg in range(13, 25)))
        for name, o in CONFIGS:
            for k in ("rx_variant", "tx_variant", "unroll", "block", "ctas_per_sm", "bulk_tile", "bulk_stages"):
                ctx.set_option(k, 0)
            for k, v in o.items():
                ctx.set_option({"v": direction + "_variant"}.get(k, k), v)
            row = []
            for lg in range(13, 25):
                n = 1 << lg
                if direction == "rx":
                    fn = lambda: ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)
                else:
                    fn = lambda: ctx.convert_tx_buffer(cf.data_ptr(), 0, out_i.data_ptr(), 0, n, 1e-6, st)
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(50):
                    fn()
                b.record()
                torch.cuda.synchronize()
                us = a.elapsed_time(b) / 50 * 1e3
                row.append(us)
                res.append(dict(direction=direction, config=name, log2_frames=lg, us=us, gbs=16 * n / us / 1e3))
            print(f"{name:22s}", " ".join(f"{u:8.2f}" for u in row), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/sweep_small.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
