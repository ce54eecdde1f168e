#!/usr/bin/env python
"""Sweep the kernel schedules of libsxgpu.so on one B200 and print achieved HBM GB/s.

Device-resident 1 GiB-per-side blocks (>> the 126 MB L2), CUDA events on the launching
stream, warm-up first.  Output: one line per configuration and a JSON file.  This is a
tuning tool; the judged numbers come from bench.py.
"""
import argparse
import itertools
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402


def time_call(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-frames", type=int, default=27)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--bulk-only", action="store_true")
    args = ap.parse_args()

    n = 1 << args.log2_frames
    ctx = Context(0)
    # A non-default torch stream: handle 0 would mean "the context's own stream" to the C ABI
    # and the CUDA events below would bracket nothing.
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    st = side.cuda_stream
    assert st != 0
    i2s = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    cf = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    out_i = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    ctx.synth_frames(i2s.data_ptr(), 0, n, 0x53581255, st)
    ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)
    torch.cuda.synchronize()

    results = []

    def run(name, fn, bytes_per_frame, **opts):
        for k in ("rx_variant", "tx_variant", "unroll", "block", "ctas_per_sm", "bulk_tile", "bulk_stages"):
            ctx.set_option(k, 0)
        for k, v in opts.items():
            ctx.set_option(k, v)
        t = time_call(fn, args.iters)
        gbs = n * bytes_per_frame / t / 1e9
        rec = dict(kernel=name, opts=opts, ms=t * 1e3, gbs=gbs, gsps=n / t / 1e9)
        results.append(rec)
        print(f"{name:10s} {json.dumps(opts):70s} {t*1e3:8.3f} ms {gbs:8.1f} GB/s", flush=True)

    # baseline: torch's own device copy of the same bytes (what MEASURED_PEAKS.json measures)
    t = time_call(lambda: out_i.copy_(i2s), args.iters)
    print(f"torch copy_ 1 GiB: {t*1e3:.3f} ms {2*8*n/t/1e9:.1f} GB/s", flush=True)
    results.append(dict(kernel="torch_copy", opts={}, ms=t * 1e3, gbs=2 * 8 * n / t / 1e9))

    rx = lambda: ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)
    tx = lambda: ctx.convert_tx_buffer(cf.data_ptr(), 0, out_i.data_ptr(), 0, n, 1e-6, st)

    vec_space = list(itertools.product((1, 2), (2, 4, 8), (256, 512), (0, 2, 4, 8)))
    bulk_space = list(itertools.product(((4096, 3), (3072, 4), (2048, 6), (2048, 5), (2048, 4), (2048, 3), (1024, 6), (1024, 4)), (256, 512), (0, 2)))
    if args.bulk_only:
        vec_space = []
    if args.quick:
        vec_space = [(1, 4, 256, 0), (2, 4, 256, 0), (2, 8, 256, 0), (2, 2, 512, 0)]
        bulk_space = [((2048, 4), 256, 0), ((1024, 4), 256, 0)]
    for name, fn, key in (("rx_cf32", rx, "rx_variant"), ("tx_cf32", tx, "tx_variant")):
        for variant, unroll, block, cps in vec_space:
            run(name, fn, 16, **{key: variant, "unroll": unroll, "block": block, "ctas_per_sm": cps})
        for (tile, stages), block, cps in bulk_space:
            run(name, fn, 16, **{key: 3, "bulk_tile": tile, "bulk_stages": stages, "block": block, "ctas_per_sm": cps})

    lb = lambda: ctx.convert_loopback(i2s.data_ptr(), cf.data_ptr(), out_i.data_ptr(), n, 1e-6, st)
    run("loopback24", lb, 24)
    lb2 = lambda: ctx.convert_loopback(i2s.data_ptr(), None, out_i.data_ptr(), n, 1e-6, st)
    run("loopback16", lb2, 16)

    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(results, indent=1))
    best = {}
    for r in results:
        if r["kernel"] not in best or r["gbs"] > best[r["kernel"]]["gbs"]:
            best[r["kernel"]] = r
    print("BEST:")
    for k, r in best.items():
        print(f"  {k}: {r['gbs']:.1f} GB/s {r['opts']}")


if __name__ == "__main__":
    main()
