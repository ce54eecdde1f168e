#!/usr/bin/env python
"""Which kernel schedule drives PCIe best when the kernel itself reads and writes pinned host memory?"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402

ctx = Context(0)
nmax = 1 << 24
src = torch.from_numpy(np.random.default_rng(1).integers(-2**31, 2**31, size=2 * nmax, dtype=np.int64).astype(np.int32)).pin_memory()
dst = torch.empty(2 * nmax, dtype=torch.float32).pin_memory()
ctx.set_option("host_mode", 2)
res = []
for lg in (12, 16, 18, 20, 22, 24):
    n = 1 << lg
    row = {}
    for name, opts in (("vec128", dict(zero_copy_variant=1)), ("vec128 u8", dict(zero_copy_variant=1, unroll=8)),
                       ("vec256 u4", dict(zero_copy_variant=2)), ("vec256 u8 b512", dict(zero_copy_variant=2, unroll=8, block=512)),
                       ("bulk 1024x4", dict(zero_copy_variant=3, bulk_tile=1024, bulk_stages=4)),
                       ("bulk 2048x4", dict(zero_copy_variant=3, bulk_tile=2048, bulk_stages=4)),
                       ("bulk 512x4 c4", dict(zero_copy_variant=3, bulk_tile=512, bulk_stages=4, block=128, ctas_per_sm=4))):
        for k in ("unroll", "block", "ctas_per_sm", "bulk_tile", "bulk_stages"):
            ctx.set_option(k, 0)
        for k, v in opts.items():
            ctx.set_option(k, v)
        reps = max(3, min(200, int(1e8 // n)))
        for _ in range(2):
            ctx.convert_rx_buffer_host(src.data_ptr(), 0, dst.data_ptr(), 0, n)
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.convert_rx_buffer_host(src.data_ptr(), 0, dst.data_ptr(), 0, n)
        t = (time.perf_counter() - t0) / reps
        row[name] = 8 * n / t / 1e9
        res.append(dict(log2_frames=lg, schedule=name, us=t * 1e6, gbs_each_way=8 * n / t / 1e9))
    print(f"2^{lg:2d}: " + "  ".join(f"{k} {v:5.1f}" for k, v in row.items()) + "  GB/s each way", flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sweep_zero_copy.json").write_text(json.dumps(res, indent=1))
