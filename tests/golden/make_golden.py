#!/usr/bin/env python
"""Regenerate the golden fixtures from the UNMODIFIED reference driver.

Run in the build container, where /root/reference exists:

    make -C oracle && python tests/golden/make_golden.py

It loads oracle/_ref/libsx_ref.so (the reference's SoapySX.cpp compiled from where it lies
against this repo's SoapySDR/ALSA stand-ins) and writes

  tests/golden/convert_kat.json    known-answer vectors of convert_rx_buffer / convert_tx_buffer
                                   (SoapySX.cpp:103-137) -- hex words in, hex words out
  tests/golden/stream_traces.json  what an application observes when the scenarios in
                                   tests/sxstream.py drive the reference's readStream/writeStream
                                   (SoapySX.cpp:868-1105) over the deterministic ALSA stub
  tests/golden/stream_fuzz_traces.json  the same for the 40 default seeds of the random-script
                                   fuzzer in tests/test_gpu_stream_fuzz.py

The reference itself ships no golden vectors (SoapySX/test/README.md:1-4); these files are the
pin.  TX vectors are restricted to the domain where the reference is defined C++ (no NaN, both
components < 1.0); the few entries under "tx_arm_semantics" are NOT produced by this x86 build
-- they state the ARM/saturating answers the product follows there.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import sxstream  # noqa: E402
import sxtest  # noqa: E402


def hx(a):
    return [format(int(x), "08x") for x in np.ascontiguousarray(a).view(np.uint32)]


def main():
    ref = sxtest.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libsx_ref.so is missing: run `make -C oracle` where /root/reference exists")

    rx_in = np.concatenate([sxtest.rx_structured(), sxtest.rx_uniform(256, seed=sxtest.SEED)])
    kat = {"source": "oracle/_ref/libsx_ref.so (unmodified SoapySX.cpp, g++ -O3 -DNDEBUG, x86-64)",
           "rx": {"in": hx(rx_in), "out": hx(sxtest.ref_rx(ref, rx_in))}, "tx": []}
    sets = {
        "uniform": sxtest.tx_uniform(256),
        "gaussian_neg_clamp": sxtest.tx_gaussian_defined(256),
        "circle_default": sxtest.tx_threshold_circle(256, sxtest.THR2_DEFAULT),
        "circle_quarter": sxtest.tx_threshold_circle(128, 0.25),
    }
    sp = sxtest.tx_specials()
    keep = sxtest.in_defined_domain(sp)
    sets["specials_defined_domain"] = sp.reshape(-1, 2)[keep].ravel()
    for name, f in sets.items():
        for thr2 in (sxtest.THR2_DEFAULT, 0.0, 0.25):
            assert sxtest.in_defined_domain(f).all(), name
            kat["tx"].append({"set": name, "thr2": hx(np.float32(thr2))[0], "in": hx(f),
                              "out": hx(sxtest.ref_tx(ref, f, thr2))})
    # ARM / saturating semantics where the reference is undefined C++ (not from this build).
    kat["tx_arm_semantics"] = {
        "note": "reference is UB here (SoapySX.cpp:124-125); ARM fcvtzs saturates, NaN -> 0",
        "thr2": hx(np.float32(sxtest.THR2_DEFAULT))[0],
        "cases": [
            {"in": hx(np.array([1.0, 0.0], np.float32)), "out": ["7fffffff", "00000000"]},
            {"in": hx(np.array([0.0, 1.0], np.float32)), "out": ["00000003", "7ffffffc"]},
            {"in": hx(np.array([2.0, np.inf], np.float32)), "out": ["7fffffff", "7ffffffc"]},
            {"in": hx(np.array([np.nan, 0.5], np.float32)), "out": ["00000000", "40000000"]},
            {"in": hx(np.array([0.5, np.nan], np.float32)), "out": ["40000000", "00000000"]},
            {"in": hx(np.array([-np.inf, -2.0], np.float32)), "out": ["80000003", "80000000"]},
        ],
    }
    (HERE / "convert_kat.json").write_text(json.dumps(kat, indent=0, separators=(",", ":")) + "\n")

    h = sxstream.Harness(sxstream.REF_LIB)
    traces = {name: sxstream.normalise(fn(h)) for name, fn in sxstream.SCENARIOS.items()}
    (HERE / "stream_traces.json").write_text(json.dumps(traces, indent=0, separators=(",", ":")) + "\n")
    # the default seeds of the stream-level differential fuzzer (tests/test_gpu_stream_fuzz.py)
    import test_gpu_stream_fuzz as fuzz
    fuzz_traces = {str(seed): fuzz.run_script(h, fuzz.make_script(seed)) for seed in range(1000, 1040)}
    (HERE / "stream_fuzz_traces.json").write_text(json.dumps(fuzz_traces, separators=(",", ":")) + "\n")
    print("wrote", HERE / "stream_fuzz_traces.json", (HERE / "stream_fuzz_traces.json").stat().st_size, "bytes")
    print("wrote", HERE / "convert_kat.json", (HERE / "convert_kat.json").stat().st_size, "bytes")
    print("wrote", HERE / "stream_traces.json", (HERE / "stream_traces.json").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
