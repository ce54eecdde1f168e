"""CPU tests of the host-side bookkeeping rules the driver=sx device uses (csrc/host/
stream_plan.hpp, exported as sxplan_*), against the oracle's restatement of the reference's
arithmetic, and of the SoapySDR/ALSA stand-ins."""
import ctypes as C

import pytest
from hypothesis import given, settings, strategies as st

import sxstream


@pytest.fixture(scope="module")
def host():
    from sxxcvr_b200 import _build
    lib = C.CDLL(str(_build.build_soapy_module()))
    lib.sxplan_geometry.argtypes = [C.c_ulong, C.POINTER(C.c_ulong), C.POINTER(C.c_ulong)]
    lib.sxplan_geometry.restype = None
    lib.sxplan_overrun_skip.argtypes = [C.c_long, C.c_ulong, C.c_ulong]
    lib.sxplan_overrun_skip.restype = C.c_ulong
    lib.sxplan_trim_nonblocking.argtypes = [C.c_ulong, C.c_long, C.c_long]
    lib.sxplan_trim_nonblocking.restype = C.c_ulong
    lib.sxplan_place_tx_block.argtypes = [C.c_int64, C.c_long, C.c_int, C.c_int64, C.c_ulong, C.POINTER(C.c_int),
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.sxplan_place_tx_block.restype = None
    return lib


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 10**6))
def test_geometry_matches_reference_rule(host, oracle, period):
    p, b, op, ob = C.c_ulong(), C.c_ulong(), C.c_ulong(), C.c_ulong()
    host.sxplan_geometry(period, C.byref(p), C.byref(b))
    oracle.sxo_alsa_sizes(period, C.byref(op), C.byref(ob))
    assert (p.value, b.value) == (op.value, ob.value)
    assert b.value % p.value == 0 and b.value <= 65536 and p.value <= 65536


@settings(max_examples=300, deadline=None)
@given(st.integers(-10, 10**7), st.sampled_from([256, 1024, 1000, 4096, 65536, 3]))
def test_overrun_skip_matches_reference_rule(host, oracle, pending, period):
    p, b = C.c_ulong(), C.c_ulong()
    host.sxplan_geometry(period, C.byref(p), C.byref(b))
    got = host.sxplan_overrun_skip(pending, b.value, p.value)
    assert got == oracle.sxo_rx_overrun_skip(pending, b.value, p.value)
    if pending > b.value:
        assert got % p.value == 0 and got >= pending - b.value + p.value


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 10**9), st.integers(-10**6, 10**6), st.booleans(), st.integers(0, 10**9),
       st.sampled_from([256, 1024, 4096]))
def test_tx_placement_matches_reference_rule(host, oracle, position, queued, timed, ticks, period):
    d, w, j = C.c_int(), C.c_int64(), C.c_int64()
    host.sxplan_place_tx_block(position, queued, int(timed), ticks, period, C.byref(d), C.byref(w), C.byref(j))
    playing = position - queued
    if timed:
        assert w.value == ticks and bool(d.value) == (playing > ticks) and j.value == 0
    else:
        jump = oracle.sxo_tx_underrun_forward(playing, position, period)
        assert not d.value and j.value == jump and w.value == position + jump


def test_nonblocking_trim(host):
    f = host.sxplan_trim_nonblocking
    assert f(256, 100, 100000) == 256          # blocking call: untouched
    assert f(256, 100, 0) == 100
    assert f(256, 100, -1) == 100
    assert f(256, 0, 0) == 0 and f(256, -5, 0) == 0
    assert f(256, 1000, 0) == 256


def test_device_construction_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = sxstream.Harness(sxstream.PRODUCT_LIB)
    assert h.lib.sxh_enumerate(b"driver=sx") == b"driver=sx, label=sx"     # the probe never needs hardware
    with pytest.raises(sxstream.Threw) as e:
        h.device()
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("helpers", [0, 1, 3, 7])
def test_parallel_bounce_copier(tmp_path, helpers):
    """csrc/host/par_copy.hpp (the bounce copies of pageable callers in the *_host entry points):
    every size class, odd sizes and offsets, guard bytes around the destination, many copies
    through one pool -- under ThreadSanitizer when the toolchain links it."""
    import os
    import subprocess
    from sxxcvr_b200 import _build
    src = _build.ROOT / "tests" / "native" / "par_copy_test.cpp"
    exe = tmp_path / "par_copy_test"
    base = [os.environ.get("CXX", "g++"), "-std=c++17", "-g", "-Wall", "-Wextra", "-pthread",
            "-I", str(_build.CSRC), str(src), "-o", str(exe)]
    tsan = subprocess.run(base + ["-O1", "-fsanitize=thread"], capture_output=True, text=True)
    rounds = "1"
    if tsan.returncode != 0:        # no libtsan here: plain build, more rounds
        subprocess.run(base + ["-O2"], check=True, capture_output=True, text=True)
        rounds = "3"
    run = subprocess.run([str(exe), str(helpers), rounds], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "ThreadSanitizer" not in run.stderr
    assert f"{helpers} helpers" in run.stdout
