#!/usr/bin/env python
"""Measure the BASELINE.json configurations that bench.py's headline line does not cover, on one
B200 (device-resident, CUDA events on the launching stream, 3 warm-ups, 20 timed iterations,
best and median):

  config 5  RX and TX conversion swept over block sizes 1 MiB .. 4 GiB (input side)
  config 4  the stream bank: S streams x (read 256 -> write 256 @ rx time + 768 frames),
            S in {1, 64, 4096, 65536}, per-launch and replayed from a CUDA graph
  config 3  timed TX bursts into a device timeline: silence + batched burst conversion
  config 2  (extension, no reference) CS16 RX / TX at 2^27 frames

Writes JSON to gpurun_out/ and prints a table.  A tuning/reporting tool; the judged line is
bench.py's.
"""
import argparse
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Bank, Context  # noqa: E402
from sxxcvr_b200.capi import Block  # noqa: E402

PEAK = 6553.0
try:
    PEAK = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return min(ms), statistics.median(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=29)
    ap.add_argument("--max-streams", type=int, default=65536)
    ap.add_argument("--out", default="gpurun_out/configs.json")
    args = ap.parse_args()

    ctx = Context(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    st = side.cuda_stream
    out = {"peak_gbs": PEAK, "config5": [], "config4": [], "config3": [], "config2_extension": []}

    # ---- config 5 ------------------------------------------------------------------------------
    nmax = 1 << args.max_log2
    i2s = torch.empty(2 * nmax, dtype=torch.int32, device="cuda")
    cf = torch.empty(2 * nmax, dtype=torch.float32, device="cuda")
    ctx.synth_frames(i2s.data_ptr(), 0, nmax, 0x53581255, st)
    ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, nmax, st)
    out_i = torch.empty(2 * nmax, dtype=torch.int32, device="cuda")
    print(f"{'block in':>10s} {'frames':>11s} | {'RX best ms':>10s} {'GB/s':>7s} {'frac':>5s} {'Gsps':>7s} | "
          f"{'TX best ms':>10s} {'GB/s':>7s} {'frac':>5s} {'Gsps':>7s}")
    for lg in range(17, args.max_log2 + 1, 2 if args.max_log2 > 20 else 1):
        n = 1 << lg
        rb, rm = timed(lambda: ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st))
        tb, tm = timed(lambda: ctx.convert_tx_buffer(cf.data_ptr(), 0, out_i.data_ptr(), 0, n, 1e-6, st))
        rec = {"log2_frames": lg, "bytes_in": 8 * n,
               "rx_best_ms": rb, "rx_median_ms": rm, "rx_gbs": 16 * n / rb / 1e6, "rx_gbs_median": 16 * n / rm / 1e6,
               "tx_best_ms": tb, "tx_median_ms": tm, "tx_gbs": 16 * n / tb / 1e6, "tx_gbs_median": 16 * n / tm / 1e6}
        out["config5"].append(rec)
        print(f"{8*n/2**20:8.0f}Mi {n:11d} | {rb:10.4f} {rec['rx_gbs']:7.0f} {rec['rx_gbs']/PEAK:5.2f} {n/rb/1e6:7.1f} | "
              f"{tb:10.4f} {rec['tx_gbs']:7.0f} {rec['tx_gbs']/PEAK:5.2f} {n/tb/1e6:7.1f}", flush=True)

    # ---- config 2 (extension) ------------------------------------------------------------------------
    n = 1 << 27
    cs = torch.empty(2 * n, dtype=torch.int16, device="cuda")
    rb, rm = timed(lambda: ctx.convert_rx_buffer_cs16(i2s.data_ptr(), 0, cs.data_ptr(), 0, n, st))
    tb, tm = timed(lambda: ctx.convert_tx_buffer_cs16(cs.data_ptr(), 0, out_i.data_ptr(), 0, n, 1e-6, st))
    out["config2_extension"] = {"note": "CS16 has no reference implementation (SoapySX.cpp:752-753)", "frames": n,
                                "rx_best_ms": rb, "rx_gbs": 12 * n / rb / 1e6, "tx_best_ms": tb, "tx_gbs": 12 * n / tb / 1e6}
    print(f"CS16 ext  2^27 frames: RX {rb:.4f} ms {12*n/rb/1e6:.0f} GB/s ({n/rb/1e6:.1f} Gsps)  "
          f"TX {tb:.4f} ms {12*n/tb/1e6:.0f} GB/s ({n/tb/1e6:.1f} Gsps)", flush=True)
    del cs

    # ---- config 3: bursts into a timeline -----------------------------------------------------------
    nb, blen, spacing = 4096, 256, 75000          # a 256-frame burst every second at 75 kHz, 4096 seconds
    timeline_frames = nb * spacing
    if timeline_frames <= nmax:
        timeline = out_i[: 2 * timeline_frames]
        bursts = cf[: 2 * nb * blen]
        blocks = [Block(bursts.data_ptr() + 8 * blen * k, timeline.data_ptr() + 8 * (spacing * k + 750), blen, 1e-6, 0)
                  for k in range(nb)]
        arr = (Block * nb)(*blocks)
        d_blocks = torch.empty(nb * 32, dtype=torch.uint8, device="cuda")
        ctx.memcpy_h2d(d_blocks.data_ptr(), arr, nb * 32, st)
        torch.cuda.synchronize()

        def sparse():
            ctx.fill_silence(timeline.data_ptr(), 0, timeline_frames, st)
            ctx.convert_batch("tx", d_blocks.data_ptr(), on_device=True, max_length=blen, stream=st, nblocks=nb)

        b, m = timed(sparse)
        dense = lambda: ctx.convert_batch("tx", d_blocks.data_ptr(), on_device=True, max_length=blen, stream=st, nblocks=nb)
        db, dm = timed(dense)
        out["config3"] = {"bursts": nb, "burst_frames": blen, "spacing_frames": spacing,
                          "sparse_timeline_best_ms": b, "sparse_timeline_gbs_written": 8 * timeline_frames / b / 1e6,
                          "bursts_only_best_us": db * 1e3, "bursts_only_msps": nb * blen / db / 1e3}
        print(f"config 3: {nb} bursts x {blen} frames, one per {spacing} frames: silence+bursts {b:.3f} ms "
              f"({8*timeline_frames/b/1e6:.0f} GB/s written); bursts alone {db*1e3:.1f} us ({nb*blen/db/1e3:.0f} Msps)", flush=True)
    del i2s, out_i
    torch.cuda.empty_cache()

    # ---- config 4: the stream bank ----------------------------------------------------------------
    lat = 10_240_000
    for S in (1, 64, 4096, 65536):
        if S > args.max_streams:
            continue
        with Bank(ctx, S, 256, 75000.0, 0.0, 1) as bank:
            buf = cf[: S * 512]

            def it():
                bank.read(buf.data_ptr(), st)
                bank.write(buf.data_ptr(), 4, None, lat, st)

            b, m = timed(it)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                it()
            gb, gm = timed(g.replay)
            rec = {"streams": S, "frames_per_block": 256, "launches_per_iteration": 3,
                   "best_us": b * 1e3, "median_us": m * 1e3, "msps_rx_plus_tx": 2 * S * 256 / b / 1e3,
                   "graph_best_us": gb * 1e3, "graph_median_us": gm * 1e3, "graph_msps_rx_plus_tx": 2 * S * 256 / gb / 1e3,
                   "ring_bytes": S * bank.ring * 8}
            out["config4"].append(rec)
            print(f"config 4: S={S:6d} streams x 256 frames: {b*1e3:9.1f} us/iter ({rec['msps_rx_plus_tx']:10.1f} Msps RX+TX)   "
                  f"graph replay {gb*1e3:9.1f} us ({rec['graph_msps_rx_plus_tx']:10.1f} Msps)", flush=True)
            _, rxp, txp = bank.positions(st)
            assert (txp - rxp == 768).all()

    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
