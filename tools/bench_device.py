#!/usr/bin/env python
"""BASELINE config 1 at the plugin level: readStream + timed writeStream through the ALSA
stand-in, block sizes 256 / 4096 / 65536 frames, for
  * the UNMODIFIED reference driver (oracle/_ref, CPU converters on the calling thread), and
  * the product driver=sx device (CUDA converters), default and lowlatency=1, pageable and
    pin=1 caller buffers.
Wall-clock per read+write pair; both run over the same deterministic stand-in, whose own
cost (it synthesises every captured frame on the CPU) is measured separately and printed.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
import sxstream  # noqa: E402


def loop(h, dev_args, stream_args, n, iters):
    with h.device("driver=sx" + dev_args) as d:
        d.set_rate(600000.0)
        period = min(n, 65536)
        rx = d.setup(sxstream.RX, args=f"period={period}{stream_args}")
        tx = d.setup(sxstream.TX, args=f"threshold=0, period={period}{stream_args}")
        d.activate(rx), d.activate(tx)
        h.lib.sx_alsa_set_sink_limit(d.play, 1 << 16)       # keep the sink small: only timing matters here
        buf = np.zeros(2 * n, np.float32)
        lat = int(round(3 * n * 1e9 / 600000.0))
        for _ in range(5):
            r, fl, t, _ = d.read(rx, n, buf=buf)
            d.write(tx, buf, n, sxstream.HAS_TIME, t + lat)
        t0 = time.perf_counter()
        for _ in range(iters):
            r, fl, t, _ = d.read(rx, n, buf=buf)
            w = d.write(tx, buf, n, sxstream.HAS_TIME, t + lat)
            assert r == n and w == n
        return (time.perf_counter() - t0) / iters * 1e6


def main():
    from sxxcvr_b200 import _build
    _build.build_soapy_module()
    product = sxstream.Harness(sxstream.PRODUCT_LIB)
    ref = sxstream.Harness(sxstream.REF_LIB) if sxstream.REF_LIB.exists() else None
    out = []
    print(f"{'frames':>7s} {'reference':>12s} {'product':>12s} {'lowlatency':>12s} {'pin=1':>12s} {'low+pin':>12s}   us per read+write pair")
    for n, iters in ((256, 2000), (4096, 1000), (65536, 200), (1 << 20, 20)):
        row = {"frames": n}
        if ref:
            row["reference_us"] = loop(ref, "", "", n, iters)
        row["product_us"] = loop(product, "", "", n, iters)
        row["product_lowlatency_us"] = loop(product, ", lowlatency=1", "", n, iters)
        row["product_pin_us"] = loop(product, "", ", pin=1", n, iters)
        row["product_lowlatency_pin_us"] = loop(product, ", lowlatency=1", ", pin=1", n, iters)
        out.append(row)
        print(f"{n:7d} {row.get('reference_us', float('nan')):12.1f} {row['product_us']:12.1f} {row['product_lowlatency_us']:12.1f} "
              f"{row['product_pin_us']:12.1f} {row['product_lowlatency_pin_us']:12.1f}", flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/bench_device.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
