#!/usr/bin/env python
"""Host-buffer entry points with PAGEABLE caller buffers (numpy arrays, what the reference's
Python callers hand to readStream/writeStream): throughput against the number of threads that
share the bounce copies (option bounce_threads), next to the same call on pinned buffers.

    python tools/bench_pageable.py --out gpurun_out/bench_pageable.json
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/bench_pageable.json")
    ap.add_argument("--log2", type=int, nargs="*", default=[20, 22, 24, 26])
    args = ap.parse_args()
    ctx = Context(0)
    nmax = 1 << max(args.log2)
    rng = np.random.default_rng(1)
    src = rng.integers(-2**31, 2**31, size=2 * nmax, dtype=np.int64).astype(np.int32)
    dst = np.zeros(2 * nmax, np.float32)
    pin_src = torch.from_numpy(src).pin_memory()
    pin_dst = torch.empty(2 * nmax, dtype=torch.float32).pin_memory()
    out = {"host_cpus": os.cpu_count(), "points": []}

    def timed(fn, n):
        fn()
        fn()
        reps = max(3, min(50, int(3e8 // n)))
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    for lg in args.log2:
        n = 1 << lg
        t = timed(lambda: ctx.convert_rx_buffer_host(pin_src.data_ptr(), 0, pin_dst.data_ptr(), 0, n), n)
        out["points"].append({"frames": n, "buffers": "pinned", "us": t * 1e6, "gbs_each_way": 8 * n / t / 1e9})
        print(f"2^{lg} frames pinned:                    {t*1e6:10.1f} us  {8*n/t/1e9:6.1f} GB/s each way", flush=True)
        for threads in (1, 2, 4, 8, 0):
            ctx.set_option("bounce_threads", threads)
            t = timed(lambda: ctx.convert_rx_buffer_host(src.ctypes.data, 0, dst.ctypes.data, 0, n), n)
            out["points"].append({"frames": n, "buffers": "pageable", "bounce_threads": threads, "us": t * 1e6,
                                  "gbs_each_way": 8 * n / t / 1e9})
            print(f"2^{lg} frames pageable, bounce_threads={threads}: {t*1e6:10.1f} us  {8*n/t/1e9:6.1f} GB/s each way",
                  flush=True)
    ctx.set_option("bounce_threads", 0)
    ctx.close()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
