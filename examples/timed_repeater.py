#!/usr/bin/env python3
"""Timed repeater over the SoapySDR Python API: every received block goes back out a fixed
latency later, addressed by its receive timestamp (the pattern of the reference's
example/linear_repeater.py, without the filters).

With a stock SoapySDR installation and the driver module installed this runs as it is.
Without one, put sxxcvr_b200/compat on PYTHONPATH:

    PYTHONPATH=sxxcvr_b200/compat python examples/timed_repeater.py --blocks 200

Prints one JSON line: blocks moved, the timestamp step, and a CRC of everything received.
"""
import argparse
import json
import zlib

import numpy as np
import SoapySDR
from SoapySDR import SOAPY_SDR_RX, SOAPY_SDR_TX, SOAPY_SDR_CF32, SOAPY_SDR_HAS_TIME


def run(blocks=100, block=256, rate=75000.0, latency=2048, threshold="0", device_args=None, pin=False):
    args = {"driver": "sx"}
    args.update(device_args or {})
    dev = SoapySDR.Device(args)
    dev.setSampleRate(SOAPY_SDR_RX, 0, rate)
    dev.setSampleRate(SOAPY_SDR_TX, 0, rate)
    dev.setFrequency(SOAPY_SDR_RX, 0, 433.9e6)
    dev.setFrequency(SOAPY_SDR_TX, 0, 433.9e6)
    stream_args = {"pin": "1"} if pin else {}
    rx = dev.setupStream(SOAPY_SDR_RX, SOAPY_SDR_CF32, [0], dict(stream_args))
    tx = dev.setupStream(SOAPY_SDR_TX, SOAPY_SDR_CF32, [0], dict(stream_args, threshold=threshold))
    dev.activateStream(rx)
    dev.activateStream(tx)

    latency_ns = SoapySDR.ticksToTimeNs(latency, rate)
    buf = np.zeros(block, dtype=np.complex64)
    crc, moved, times, tx_rets = 0, 0, [], []
    for _ in range(blocks):
        r = dev.readStream(rx, [buf], block)
        if r.ret != block or not (r.flags & SOAPY_SDR_HAS_TIME):
            raise RuntimeError("readStream: %s" % r)
        times.append(r.timeNs)
        crc = zlib.crc32(buf.view(np.uint8), crc)
        w = dev.writeStream(tx, [buf], block, SOAPY_SDR_HAS_TIME, r.timeNs + latency_ns)
        tx_rets.append(w.ret)
        moved += 1
    hw_time = dev.getHardwareTime()
    dev.deactivateStream(rx)
    dev.deactivateStream(tx)
    dev.closeStream(rx)
    dev.closeStream(tx)
    steps = sorted(set(np.diff(np.asarray(times, dtype=np.int64)).tolist()))
    return {"blocks": moved, "block": block, "first_time_ns": times[0], "time_steps_ns": steps,
            "tx_rets": sorted(set(tx_rets)), "hardware_time_ns": hw_time, "rx_crc32": crc}


def main():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--blocks", type=int, default=100)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--rate", type=float, default=75000.0)
    ap.add_argument("--latency", type=int, default=2048, help="frames between receive and transmit")
    ap.add_argument("--threshold", default="0")
    ap.add_argument("--pin", action="store_true", help="page-lock the sample buffer (stream argument pin=1)")
    a = ap.parse_args()
    print(json.dumps(run(a.blocks, a.block, a.rate, a.latency, a.threshold, pin=a.pin)))


if __name__ == "__main__":
    main()
