#!/usr/bin/env python
"""Tune the host-buffer entry points (sxgpu_convert_*_buffer_host) on one B200: copy-engine
pipeline chunk size versus the zero-copy kernel, pinned versus pageable callers, across block
sizes from one ALSA period (256 frames) to 2^26 frames.  Wall-clock around the synchronous call.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402


def bench(fn, reps, warmup=2):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=26)
    ap.add_argument("--out", default="gpurun_out/sweep_host.json")
    args = ap.parse_args()
    ctx = Context(0)
    nmax = 1 << args.max_log2
    rng = np.random.default_rng(1)
    src_np = rng.integers(-2**31, 2**31, size=2 * nmax, dtype=np.int64).astype(np.int32)
    pin_src = torch.from_numpy(src_np).pin_memory()
    pin_dst = torch.empty(2 * nmax, dtype=torch.float32).pin_memory()
    page_src = torch.from_numpy(src_np)
    page_dst = torch.empty(2 * nmax, dtype=torch.float32)
    results = []

    def run(label, n, s, d, **opts):
        for k, v in opts.items():
            ctx.set_option(k, v)
        reps = max(3, min(300, int(2e8 // max(n, 1 << 16))))
        t = bench(lambda: ctx.convert_rx_buffer_host(s.data_ptr(), 0, d.data_ptr(), 0, n), reps)
        rec = dict(label=label, frames=n, opts=opts, us=t * 1e6, msps=n / t / 1e6, gbs_each_way=8 * n / t / 1e9)
        results.append(rec)
        print(f"{label:10s} n=2^{int(np.log2(n)):2d} {json.dumps(opts):60s} {t*1e6:12.1f} us {n/t/1e6:10.1f} Msps "
              f"{8*n/t/1e9:6.1f} GB/s each way", flush=True)

    for lg in (8, 12, 14, 15, 16, 17, 18, 20):
        n = 1 << lg
        run("pinned", n, pin_src, pin_dst, host_mode=2)
        run("pinned", n, pin_src, pin_dst, host_mode=1, host_chunk_frames=max(1024, n // 4))
        run("pinned", n, pin_src, pin_dst, host_mode=1, host_chunk_frames=n)
    for lg in (22, 24, args.max_log2):
        n = 1 << lg
        run("pinned", n, pin_src, pin_dst, host_mode=2)
        run("pinned", n, pin_src, pin_dst, host_mode=0, host_chunk_frames=0)      # the defaults
        run("pinned", n, pin_src, pin_dst, host_mode=1, host_chunk_frames=0)      # ramped schedule, forced CE
        for chunk_lg in (16, 17, 18, 19, 20, 21, 22):
            run("pinned", n, pin_src, pin_dst, host_mode=1, host_chunk_frames=1 << chunk_lg)
    for lg in (8, 16, 20, 24):
        n = 1 << lg
        for chunk_lg in (17, 19, 21):
            run("pageable", n, page_src, page_dst, host_mode=1, host_chunk_frames=1 << chunk_lg)

    # reference points: raw cudaMemcpy both ways, back to back and overlapped
    n = 1 << 24
    dev = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d_only():
        dev.copy_(pin_src[: 2 * n], non_blocking=True)
        torch.cuda.synchronize()

    def both():
        with torch.cuda.stream(s1):
            dev.copy_(pin_src[: 2 * n], non_blocking=True)
        with torch.cuda.stream(s2):
            pin_dst[: 2 * n].view(torch.int32).copy_(dev, non_blocking=True)
        torch.cuda.synchronize()

    t = bench(h2d_only, 10)
    print(f"raw H2D 128 MiB: {8*n/t/1e9:.1f} GB/s")
    results.append(dict(label="raw_h2d", gbs=8 * n / t / 1e9))
    t = bench(both, 10)
    print(f"raw H2D + D2H overlapped, 128 MiB each: {8*n/t/1e9:.1f} GB/s each way")
    results.append(dict(label="raw_bidir", gbs_each_way=8 * n / t / 1e9))

    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
