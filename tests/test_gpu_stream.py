"""-m gpu: the product driver=sx device (CUDA converters behind the SoapySDR surface) must be
indistinguishable from the unmodified reference driver: same return codes, flags, timestamps,
sample bytes and playback timeline in every scenario."""
import json

import numpy as np
import pytest

import sxstream
import sxtest

pytestmark = pytest.mark.gpu

GOLDEN = json.loads(sxstream.GOLDEN_TRACES.read_text())


@pytest.fixture(scope="module")
def product():
    from sxxcvr_b200 import _build
    _build.build_soapy_module()
    return sxstream.Harness(sxstream.PRODUCT_LIB)


@pytest.mark.parametrize("name", sorted(sxstream.SCENARIOS))
def test_product_matches_golden_trace(product, name):
    got = sxstream.normalise(sxstream.SCENARIOS[name](product))
    want = GOLDEN[name]
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"{name}: step {i} differs"
    assert len(got) == len(want)


@pytest.mark.parametrize("name", sorted(sxstream.SCENARIOS))
def test_product_matches_reference_live(product, name):
    if not sxstream.REF_LIB.exists():
        pytest.skip("oracle/_ref/libsx_ref.so not present")
    ref = sxstream.Harness(sxstream.REF_LIB)
    assert sxstream.normalise(sxstream.SCENARIOS[name](product)) == sxstream.normalise(sxstream.SCENARIOS[name](ref))


def test_full_scale_burst_follows_arm_semantics(product):
    """SoapySX/test/test_timestamps.py:34 transmits np.ones: exactly +1.0, where the reference is
    undefined C++.  The product gives what the reference gives on its ARM target."""
    with product.device() as d:
        d.set_rate(75000.0)
        rx, tx = d.setup(sxstream.RX), d.setup(sxstream.TX)
        d.activate(rx), d.activate(tx)
        _, _, t, _ = d.read(rx, 256)
        ones = np.zeros(512, np.float32)
        ones[0::2] = 1.0
        assert d.write(tx, ones, 256, sxstream.HAS_TIME, t + 10_000_000) == 256
        words = d.sink(750, 256).view(np.uint32).reshape(-1, 2)
        assert (words[:, 0] == 0x7FFFFFFF).all() and (words[:, 1] == 0).all()
        assert not d.sink_written_mask(0, 750).any()                # silence before the burst


def test_read_values_are_the_oracle_of_the_synthetic_frames(product, oracle):
    with product.device() as d:
        d.set_rate(300000.0)
        rx = d.setup(sxstream.RX, args="period=4096")
        d.activate(rx)
        for k, n in enumerate((4096, 1, 65536, 100000, 255)):
            before = d.pointers()[1]
            r, fl, t, buf = d.read(rx, n)
            assert r == n and fl == sxstream.HAS_TIME
            want = sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, before, n))
            assert np.array_equal(buf.view(np.uint32), want.view(np.uint32)), (k, n)


@pytest.mark.parametrize("dev_args", ["driver=sx", "driver=sx, overlap=0"])
def test_long_reads_convert_under_the_read(product, oracle, dev_args):
    """Blocking reads of 2^22 frames and more are read piece by piece with the conversion running
    behind the read (sxgpu_convert_rx_buffer_host_gated): same values, timestamps, counters and
    return codes as the one-piece path (overlap=0) and as the reference driver."""
    sizes = ((1 << 22) + 3, 1 << 22, 3 * (1 << 21) + 1, 70000)
    ref = sxstream.Harness(sxstream.REF_LIB) if sxstream.REF_LIB.exists() else None
    with product.device(dev_args) as d:
        d.set_rate(300000.0)
        rx = d.setup(sxstream.RX, args="period=4096")
        d.activate(rx)
        got = []
        for n in sizes:
            before = d.pointers()[1]
            r, fl, t, buf = d.read(rx, n)
            assert r == n and fl == sxstream.HAS_TIME
            want = sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, before, n))
            assert np.array_equal(buf.view(np.uint32), want.view(np.uint32)), n
            got.append((r, fl, t, d.pointers()))
    if ref is not None:
        with ref.device() as d:
            d.set_rate(300000.0)
            rx = d.setup(sxstream.RX, args="period=4096")
            d.activate(rx)
            for n, g in zip(sizes, got):
                r, fl, t, _ = d.read(rx, n)
                assert (r, fl, t, d.pointers()) == g, n


def test_long_read_that_ends_in_an_error_returns_what_was_read(product, oracle):
    """A capture error injected into a later piece of a long read: the frames read so far come back
    converted (as a failing snd_pcm_readi returns them), the error on the next call."""
    n = (1 << 22) + 5
    with product.device() as d:
        d.set_rate(300000.0)
        rx = d.setup(sxstream.RX, args="period=4096")
        d.activate(rx)
        before = d.pointers()[1]
        d.inject(True, sxstream.OP_READI, -32, 2)            # the third snd_pcm_readi of the stream fails with -EPIPE
        r, fl, t, buf = d.read(rx, n)
        assert r == 2 << 20 and fl == sxstream.HAS_TIME      # two pieces of 2^20 frames landed
        want = sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, before, r))
        assert np.array_equal(buf[: 2 * r].view(np.uint32), want.view(np.uint32))


def test_capture_table_known_answers_through_readstream(product):
    kat = {0: 0x00000000, 1: 0x30000000, -1: 0xB0000000, 2**31 - 1: 0x3F800000, -2**31: 0xBF800000,
           0x7FFFFF80: 0x3F7FFFFF, 0x7FFFFFC0: 0x3F800000, 0x12345678: 0x3E11A2B4}
    words = np.array(list(kat), dtype=np.int64).astype(np.int32)
    with product.device() as d:
        rx = d.setup(sxstream.RX)
        d.capture_table(words)
        d.activate(rx)
        r, fl, t, buf = d.read(rx, 4)
        assert r == 4 and [int(x) for x in buf.view(np.uint32)] == list(kat.values())


def test_two_devices_are_independent(product, oracle):
    with product.device() as a, product.device() as b:
        ra, rb = a.setup(sxstream.RX), b.setup(sxstream.RX)
        product.lib.sx_alsa_set_capture_seed(b.cap, 99)
        a.activate(ra), b.activate(rb)
        _, _, ta, bufa = a.read(ra, 512)
        _, _, tb1, _ = b.read(rb, 256)
        _, _, tb2, bufb = b.read(rb, 256)
        assert (ta, tb1) == (0, 0) and tb2 > 0
        assert np.array_equal(bufa.view(np.uint32), sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, 0, 512)).view(np.uint32))
        assert np.array_equal(bufb.view(np.uint32),
                              sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, 256, 256, seed=99)).view(np.uint32))


def test_rx_and_tx_threads_run_concurrently(product):
    """example/plot_rxtx_response.py:65-77 runs TX on its own thread; one mutex per direction."""
    import threading
    with product.device() as d:
        d.set_rate(300000.0)
        rx, tx = d.setup(sxstream.RX), d.setup(sxstream.TX, args="threshold=0")
        d.activate(rx), d.activate(tx)
        errors = []

        def tx_loop():
            blk = sxtest.tx_uniform(1024, seed=5)
            for _ in range(200):
                if d.write(tx, blk, 1024) != 1024:
                    errors.append("short write")

        th = threading.Thread(target=tx_loop)
        th.start()
        last = -1
        for _ in range(200):
            r, fl, t, _ = d.read(rx, 1024)
            if r != 1024 or t <= last:
                errors.append(("rx", r, t, last))
            last = t
        th.join()
        assert not errors


def test_readstream_into_a_device_buffer_and_writestream_from_one(product, oracle):
    """The caller's buffs[0] may be device memory (a torch tensor): same values, no D2H/H2D."""
    import torch
    with product.device() as d:
        d.set_rate(300000.0)
        rx, tx = d.setup(sxstream.RX), d.setup(sxstream.TX, args="threshold=0")
        d.activate(rx), d.activate(tx)
        n = 4096
        dev = torch.zeros(2 * n, dtype=torch.float32, device="cuda")
        import ctypes as C
        flags, t = C.c_int(-1), C.c_longlong(-1)
        ret = product.lib.sxh_read(d.p, rx, dev.data_ptr(), n, C.byref(flags), C.byref(t), 100000)
        assert (ret, flags.value, t.value) == (n, sxstream.HAS_TIME, 0)
        want = sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, 0, n))
        assert np.array_equal(dev.cpu().numpy().view(np.uint32), want.view(np.uint32))
        f = C.c_int(sxstream.HAS_TIME)
        when = t.value + 20_000_000
        assert product.lib.sxh_write(d.p, tx, dev.data_ptr(), n, C.byref(f), when, 100000) == n
        pos = product.lib.sxh_time_ns_to_ticks(when, 300000.0)
        assert np.array_equal(d.sink(pos, n), sxtest.oracle_tx(oracle, want, 0.0))


def test_cs16_stream_format_is_an_opt_in_extension(product, oracle):
    """EXTENSION, no reference behaviour: BASELINE config 2 names a CS16 readStream path, the
    reference has none (SoapySX.cpp:752-753).  Off by default; with cs16=1 the values follow our
    own specification in oracle/sx_oracle.c."""
    import ctypes as C
    with product.device() as d:
        assert product.lib.sxh_stream_formats(d.p, sxstream.RX) == b"CF32"
        with pytest.raises(sxstream.Threw):
            d.setup(sxstream.RX, "CS16")
    with product.device("driver=sx, cs16=1") as d:
        assert product.lib.sxh_stream_formats(d.p, sxstream.RX) == b"CF32,CS16"
        d.set_rate(75000.0)
        rx, tx = d.setup(sxstream.RX, "CS16"), d.setup(sxstream.TX, "CS16", "threshold=0.25")
        d.activate(rx), d.activate(tx)
        for n in (256, 4096, 70001):
            first = d.pointers()[1]
            buf = np.zeros(2 * n, np.int16)
            flags, t = C.c_int(-1), C.c_longlong(-1)
            ret = product.lib.sxh_read(d.p, rx, buf.ctypes.data, n, C.byref(flags), C.byref(t), 100000)
            assert ret == n and flags.value == sxstream.HAS_TIME
            assert np.array_equal(buf, sxtest.oracle_rx_cs16(oracle, sxtest.synth_frames(oracle, first, n)))
            f = C.c_int(sxstream.HAS_TIME)
            when = t.value + 1_000_000_000
            assert product.lib.sxh_write(d.p, tx, buf.ctypes.data, n, C.byref(f), when, 100000) == n
            pos = product.lib.sxh_time_ns_to_ticks(when, 75000.0)
            thr2 = float(np.float32(0.25) * np.float32(0.25))
            assert np.array_equal(d.sink(pos, n), sxtest.oracle_tx_cs16(oracle, buf, thr2))


@pytest.mark.parametrize("args", [["100"], ["50", "1024", "300000"], ["30", "1000", "20833.333333333332", "clock=32e6"]])
def test_cpp_repeater_example(args):
    """examples/repeater.cpp: a C++ application written against the SoapySDR API only."""
    import subprocess
    from sxxcvr_b200 import _build
    exe = _build.build_examples()
    p = subprocess.run([str(exe)] + args, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.startswith("OK"), (p.stdout, p.stderr[-500:])


def test_lowlatency_device_argument_matches_golden_traces(product):
    """lowlatency=1 routes period-sized blocks through the resident converter; observable
    behaviour must not change."""
    for name in ("repeater", "timed_bursts", "nonblocking"):
        class Low:
            lib = product.lib

            @staticmethod
            def device(args="driver=sx"):
                return product.device(args + ", lowlatency=1")
        got = sxstream.normalise(sxstream.SCENARIOS[name](Low))
        assert got == GOLDEN[name], name


def test_pin_stream_argument_page_locks_reused_caller_buffers(product, oracle):
    """pin=1 (ours): a pageable buffer the application reuses is page-locked once and then moved
    by DMA.  Observable results are unchanged; the bytes crossing PCIe are counted either way."""
    import time
    n = 1 << 19          # 4 reads of 2^19 frames plus the look-ahead stay inside the stand-in's 4 Mi-frame sink
    results = {}
    for pin in ("0", "1"):
        with product.device() as d:
            d.set_rate(600000.0)
            rx = d.setup(sxstream.RX, args=f"period=65536, pin={pin}")
            tx = d.setup(sxstream.TX, args=f"threshold=0, period=65536, pin={pin}")
            d.activate(rx), d.activate(tx)
            product.lib.sx_alsa_set_sink_limit(d.play, 1 << 24)
            buf = np.zeros(2 * n, np.float32)           # pageable, reused for every call
            stamps = []
            for k in range(3):
                t0 = time.perf_counter()
                r, fl, t, _ = d.read(rx, n, buf=buf)
                stamps.append(time.perf_counter() - t0)
                assert r == n
                # the far-future write below lets the capture side overrun, so the block's first
                # frame is whatever the timestamp says, not simply the previous end
                first = product.lib.sxh_time_ns_to_ticks(t, 600000.0)
                want = sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, first, n))
                assert np.array_equal(buf.view(np.uint32), want.view(np.uint32)), (pin, k)
                assert d.write(tx, buf, n, sxstream.HAS_TIME, t + 1_000_000_000) == n
            pos = product.lib.sxh_time_ns_to_ticks(t + 1_000_000_000, 600000.0)
            assert pos + n < (1 << 24)
            assert np.array_equal(d.sink(pos, 4096), sxtest.oracle_tx(oracle, want[:8192], 0.0))
            results[pin] = min(stamps[1:])
            d.close_stream(rx), d.close_stream(tx)
    print(f"readStream of 2^19 frames into a pageable buffer: {results['0']*1e3:.2f} ms, with pin=1 {results['1']*1e3:.2f} ms")
