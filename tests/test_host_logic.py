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
    lib.sxplan_clock_after_forward.argtypes = [C.c_int64] * 5
    lib.sxplan_clock_after_forward.restype = C.c_int64
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


@pytest.mark.parametrize("helpers", [0, 2])
def test_host_pipeline_pieces(tmp_path, helpers):
    """csrc/host/par_copy.hpp: the chunk schedule, and the owner/sidekick handshake of the *_host
    pipeline (including an aborted run) -- under ThreadSanitizer when the toolchain links it."""
    import os
    import subprocess
    from sxxcvr_b200 import _build
    src = _build.ROOT / "tests" / "native" / "pipeline_test.cpp"
    exe = tmp_path / "pipeline_test"
    base = [os.environ.get("CXX", "g++"), "-std=c++17", "-g", "-Wall", "-Wextra", "-pthread",
            "-I", str(_build.CSRC), str(src), "-o", str(exe)]
    tsan = subprocess.run(base + ["-O1", "-fsanitize=thread"], capture_output=True, text=True)
    if tsan.returncode != 0:
        subprocess.run(base + ["-O2"], check=True, capture_output=True, text=True)
    run = subprocess.run([str(exe), str(helpers)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "ThreadSanitizer" not in run.stderr
    assert "pipeline ok" in run.stdout


def forward_loop(clock, pos, target, ring, period):
    """The reference's forward-and-wait loop (SoapySX.cpp:1043-1073) over the virtual-clock model
    of the stream bank, one turn at a time: what sxplan::clock_after_forward states in closed form."""
    gap = target - pos
    while gap > 0:
        fits = max(clock + ring - pos, 0)
        if gap < fits:
            moved = gap
        else:
            moved = fits
            room_after = clock + ring - (pos + moved)
            if room_after < period:
                clock += period - room_after
        pos += moved
        gap -= moved
    return clock


@settings(max_examples=2000, deadline=None)
@given(st.integers(0, 10**7), st.integers(-70000, 200000), st.integers(0, 600000),
       st.sampled_from([1, 3, 256, 1000, 4096, 65536]))
def test_forward_clock_closed_form_matches_the_loop(host, clock, lead, gap, period):
    ring = 65536 // period * period
    pos = clock + lead                      # the write pointer leads (or trails) the clock
    target = pos + gap
    assert host.sxplan_clock_after_forward(clock, pos, target, ring, period) == forward_loop(clock, pos, target, ring, period)


def test_forward_clock_takes_a_garbage_timestamp_in_constant_time(host):
    import time
    t0 = time.perf_counter()
    c = host.sxplan_clock_after_forward(0, 0, 10**15, 65536, 256)
    assert time.perf_counter() - t0 < 0.01
    assert c == 10**15 - 65536 + 256 - (10**15 - 65536) % 256 or c > 0      # lands within a period of target - ring
    assert 0 <= (10**15 - 65536 + 256) - c < 256


def test_stub_devices_on_separate_threads_do_not_share_a_lock():
    """The ALSA stand-in locks per clock group: two devices driven from two threads make progress
    independently, and a device's RX and TX threads (one clock group) still see a consistent clock."""
    import threading
    import numpy as np
    from sxxcvr_b200 import _build
    if not sxstream.REF_LIB.exists():
        pytest.skip("needs oracle/_ref (the only driver=sx device that runs without a GPU)")
    h = sxstream.Harness(sxstream.REF_LIB)
    results = {}

    def run(k):
        with h.device() as d:
            d.set_rate(75000.0)
            h.lib.sx_alsa_set_capture_seed(d.cap, 100 + k)
            rx, tx = d.setup(sxstream.RX), d.setup(sxstream.TX, args="threshold=0")
            d.activate(rx), d.activate(tx)
            crcs = []
            for _ in range(300):
                r, fl, t, buf = d.read(rx, 256)
                w = d.write(tx, buf, 256, sxstream.HAS_TIME, t + 10_240_000)
                crcs.append((r, t, sxstream.crc(buf), w))
            results[k] = (crcs, d.pointers())

    ts = [threading.Thread(target=run, args=(k,)) for k in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    serial = {}
    for k in range(4):
        saved = dict(results)
        run(k)
        serial[k] = results[k]
        results.update(saved)
    for k in range(4):
        assert results[k] == serial[k], k
        assert results[k][1] == [300 * 256, 300 * 256, results[k][1][2], 299 * 256 + 768 + 256]
