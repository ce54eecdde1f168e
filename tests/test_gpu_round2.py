"""-m gpu: what round 2 added to the C ABI -- batched mid-size blocks on the bulk-async schedule,
the bulk loopback, flag-completed small calls, one host lane per direction, the ramped chunk
schedule, the resident converter's quit flag -- and a shorter parity chain: the CUDA entry
points against the committed golden vectors (generated from the unmodified reference) and
against the reference's own converters run live on the GPU box.  Tolerance: zero."""
import ctypes as C
import json
import threading
import time

import numpy as np
import pytest

import sxtest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def dev(arr):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def host(t):
    return t.cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def unhex(words, dtype):
    return np.array([int(w, 16) for w in words], dtype=np.uint32).view(dtype)


RESET = dict(batch_variant=0, loopback_variant=0, small_mode=0, host_mode=0, host_chunk_frames=0,
             host_chunk_min_frames=0, bounce_threads=0, resident_max_frames=0, host_in_mode=0, host_out_mode=0,
             bounce_nt=1)


@pytest.fixture(autouse=True)
def reset_options(ctx):
    yield
    for k, v in RESET.items():
        ctx.set_option(k, v)


# ---------------------------------------------------------------------------------------------
# Parity chain, one hop: CUDA against the golden vectors and against the live reference
# ---------------------------------------------------------------------------------------------
KAT = json.loads((sxtest.GOLDEN / "convert_kat.json").read_text())


def test_golden_rx_vectors_through_the_cuda_entry_point(ctx):
    """tests/golden/convert_kat.json was produced by the unmodified convert_rx_buffer
    (SoapySX.cpp:103-112); the CUDA entry point must reproduce every word."""
    words = unhex(KAT["rx"]["in"], np.int32)
    want = unhex(KAT["rx"]["out"], np.uint32)
    src = dev(words)
    dst = torch.zeros(words.size, dtype=torch.float32, device="cuda")
    ctx.convert_rx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, words.size // 2)
    ctx.stream_sync()
    assert np.array_equal(bits(host(dst)), want)
    out = np.zeros(words.size, np.float32)       # and through the host-buffer entry (readStream's call)
    ctx.convert_rx_buffer_host(words.ctypes.data, 0, out.ctypes.data, 0, words.size // 2)
    assert np.array_equal(bits(out), want)


def test_golden_tx_vectors_through_the_cuda_entry_point(ctx):
    """Every TX vector the unmodified convert_tx_buffer (SoapySX.cpp:116-137) produced on its
    defined domain, and the ARM-semantics cases of the undefined one (parity policy)."""
    assert len(KAT["tx"]) >= 15
    for case in KAT["tx"]:
        f = unhex(case["in"], np.float32)
        thr2 = float(unhex([case["thr2"]], np.float32)[0])
        want = unhex(case["out"], np.int32)
        src = dev(f)
        dst = torch.zeros(f.size, dtype=torch.int32, device="cuda")
        ctx.convert_tx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, f.size // 2, thr2)
        ctx.stream_sync()
        assert np.array_equal(host(dst), want), (case["set"], case["thr2"])
        out = np.zeros(f.size, np.int32)
        ctx.convert_tx_buffer_host(f.ctypes.data, 0, out.ctypes.data, 0, f.size // 2, thr2)
        assert np.array_equal(out, want), (case["set"], case["thr2"], "host")
    arm = KAT["tx_arm_semantics"]
    thr2 = float(unhex([arm["thr2"]], np.float32)[0])
    for case in arm["cases"]:
        f = unhex(case["in"], np.float32)
        src = dev(f)
        dst = torch.zeros(2, dtype=torch.int32, device="cuda")
        ctx.convert_tx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, 1, thr2)
        ctx.stream_sync()
        assert np.array_equal(host(dst), unhex(case["out"], np.int32)), case


@pytest.mark.parametrize("nframes", [256, 4096, 65536, (1 << 22) + 3])
def test_cuda_against_the_live_reference_converters(ctx, ref, nframes):
    """CUDA vs the unmodified converters in oracle/_ref, run here, on the inputs of BASELINE
    configs 1 and 3: full-range uniform I2S words; uniform, Gaussian with
    the negative clamp, and threshold-circle CF32 (all inside the reference's defined domain)."""
    words = sxtest.rx_uniform(nframes, seed=sxtest.SEED)
    src = dev(words)
    dst = torch.zeros(2 * nframes, dtype=torch.float32, device="cuda")
    ctx.convert_rx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, nframes)
    ctx.stream_sync()
    assert np.array_equal(bits(host(dst)), bits(sxtest.ref_rx(ref, words)))
    for name, f in (("uniform", sxtest.tx_uniform(nframes, seed=sxtest.SEED + 1)),
                    ("gaussian", sxtest.tx_gaussian_defined(nframes, seed=2)),
                    ("circle", sxtest.tx_threshold_circle(min(nframes, 65536), sxtest.THR2_DEFAULT))):
        assert sxtest.in_defined_domain(f).all()
        for thr2 in (sxtest.THR2_DEFAULT, 0.0):
            fs = dev(f)
            out = torch.zeros(f.size, dtype=torch.int32, device="cuda")
            ctx.convert_tx_buffer(fs.data_ptr(), 0, out.data_ptr(), 0, f.size // 2, thr2)
            ctx.stream_sync()
            assert np.array_equal(host(out), sxtest.ref_tx(ref, f, thr2)), (name, thr2)


# ---------------------------------------------------------------------------------------------
# Batched mid-size blocks (BASELINE config 5's small end) in one launch
# ---------------------------------------------------------------------------------------------
def make_blocks(src_ptr, dst_ptr, lengths, offsets, thr):
    from sxxcvr_b200.capi import Block
    return [Block(src_ptr + 8 * o, dst_ptr + 8 * o, n, t, 0) for n, o, t in zip(lengths, offsets, thr)]


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", ["64x1MiB", "16x4MiB", "4x16MiB", "ragged", "misaligned", "device_list"])
def test_batched_mid_size_blocks(ctx, oracle, variant, shape):
    rng = np.random.default_rng(11)
    if shape == "64x1MiB":
        lengths = [1 << 17] * 64
    elif shape == "16x4MiB":
        lengths = [1 << 19] * 16
    elif shape == "4x16MiB":
        lengths = [1 << 21] * 4
    elif shape == "ragged":      # odd lengths, empty blocks, one frame, just over / under a tile
        lengths = [70001, 0, 1, 2047, 2048, 2049, 4097, 0, 131071, 5000, 3, 262145]
    elif shape == "misaligned":  # blocks that start 8 bytes off a 16-byte boundary (direct tiles) among aligned ones
        lengths = [8192, 8193, 70000, 4100, 9000, 16384]
    else:
        lengths = [1 << 16] * 8 + [12345, 70001]
    gaps = [3 if shape == "misaligned" and i % 2 else (rng.integers(0, 4) * 2 if shape != "misaligned" else 2)
            for i in range(len(lengths))]
    offsets, at = [], 0
    for n, g in zip(lengths, gaps):
        at += int(g)
        offsets.append(at)
        at += n
    total = at + 4
    thr = [sxtest.THR2_DEFAULT if i % 2 else 0.0 for i in range(len(lengths))]
    ctx.set_option("batch_variant", variant)

    words = sxtest.rx_uniform(total, seed=31)
    src = dev(words)
    dst = torch.zeros(2 * total, dtype=torch.float32, device="cuda")
    blocks = make_blocks(src.data_ptr(), dst.data_ptr(), lengths, offsets, thr)
    if shape == "device_list":
        from sxxcvr_b200.capi import Block
        arr = (Block * len(blocks))(*blocks)
        d_list = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).cuda()
        ctx.convert_batch("rx", d_list.data_ptr(), on_device=True, max_length=max(lengths), nblocks=len(blocks))
    else:
        ctx.convert_batch("rx", blocks)
    ctx.stream_sync()
    got = host(dst)
    want_all = sxtest.oracle_rx(oracle, words)
    covered = np.zeros(2 * total, bool)
    for n, o in zip(lengths, offsets):
        assert np.array_equal(bits(got[2 * o: 2 * (o + n)]), bits(want_all[2 * o: 2 * (o + n)])), (shape, n, o)
        covered[2 * o: 2 * (o + n)] = True
    assert (got[~covered] == 0).all()           # nothing outside the blocks was touched

    f = sxtest.tx_uniform(total, seed=32)
    fsrc = dev(f)
    idst = torch.zeros(2 * total, dtype=torch.int32, device="cuda")
    tblocks = make_blocks(fsrc.data_ptr(), idst.data_ptr(), lengths, offsets, thr)
    ctx.convert_batch("tx", tblocks)
    ctx.stream_sync()
    goti = host(idst)
    for n, o, t in zip(lengths, offsets, thr):
        assert np.array_equal(goti[2 * o: 2 * (o + n)], sxtest.oracle_tx(oracle, f[2 * o: 2 * (o + n)], t)), (shape, n, o)
    assert (goti[~covered] == 0).all()


@pytest.mark.parametrize("variant", [0, 2, 3])
def test_batched_device_list_with_an_understated_max_length(ctx, oracle, variant):
    """max_length only sizes the grid: a device-resident list whose longest block is longer than the
    caller said is still converted whole (the last CTA of a block takes what is left of it)."""
    from sxxcvr_b200.capi import Block
    lengths = [70001, 5000, 8192, 123457, 6]
    offsets, at = [], 2
    for n in lengths:
        offsets.append(at)
        at += n + 2
    total = at
    words = sxtest.rx_uniform(total, seed=77)
    src = dev(words)
    dst = torch.zeros(2 * total, dtype=torch.float32, device="cuda")
    blocks = make_blocks(src.data_ptr(), dst.data_ptr(), lengths, offsets, [0.0] * len(lengths))
    arr = (Block * len(blocks))(*blocks)
    d_list = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).cuda()
    ctx.set_option("batch_variant", variant)
    ctx.convert_batch("rx", d_list.data_ptr(), on_device=True, max_length=8192, nblocks=len(blocks))
    ctx.stream_sync()
    ctx.set_option("batch_variant", 0)
    got, want = host(dst), sxtest.oracle_rx(oracle, words)
    covered = np.zeros(2 * total, bool)
    for n, o in zip(lengths, offsets):
        assert np.array_equal(bits(got[2 * o: 2 * (o + n)]), bits(want[2 * o: 2 * (o + n)])), (n, o)
        covered[2 * o: 2 * (o + n)] = True
    assert (got[~covered] == 0).all()


def test_batched_blocks_in_place_and_back_to_back(ctx, oracle):
    """In place (src == dest, allowed for the equal-width conversions) and several batches queued
    back to back on one stream, each with its own host-resident descriptor list."""
    n, nb = 1 << 15, 24
    words = sxtest.rx_uniform(n * nb, seed=5)
    buf = dev(words)
    from sxxcvr_b200.capi import Block
    for part in range(3):
        blocks = [Block(buf.data_ptr() + 8 * n * b, buf.data_ptr() + 8 * n * b, n, 0.0, 0)
                  for b in range(part * 8, part * 8 + 8)]
        ctx.convert_batch("rx", blocks)
    ctx.stream_sync()
    assert np.array_equal(bits(host(buf)), bits(sxtest.oracle_rx(oracle, words)))


# ---------------------------------------------------------------------------------------------
# Fused loopback on the bulk-async schedule
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [1, 2, 3, 2048, 2049, 4096 * 148 + 2, (1 << 22) + 1])
def test_loopback_schedules(ctx, oracle, variant, n):
    ctx.set_option("loopback_variant", variant)
    words = sxtest.rx_uniform(n, seed=n)
    src = dev(words)
    mid = torch.zeros(2 * n + 4, dtype=torch.float32, device="cuda")
    out = torch.zeros(2 * n + 4, dtype=torch.int32, device="cuda")
    ctx.convert_loopback(src.data_ptr(), mid.data_ptr(), out.data_ptr(), n, sxtest.THR2_DEFAULT)
    ctx.stream_sync()
    want_mid = sxtest.oracle_rx(oracle, words)
    want_out = sxtest.oracle_tx(oracle, want_mid, sxtest.THR2_DEFAULT)
    assert np.array_equal(bits(host(mid)[:2 * n]), bits(want_mid))
    assert np.array_equal(host(out)[:2 * n], want_out)
    assert not host(mid)[2 * n:].any() and not host(out)[2 * n:].any()
    out2 = torch.zeros(2 * n, dtype=torch.int32, device="cuda")
    ctx.convert_loopback(src.data_ptr(), None, out2.data_ptr(), n, sxtest.THR2_DEFAULT)
    ctx.stream_sync()
    assert np.array_equal(host(out2), want_out)


# ---------------------------------------------------------------------------------------------
# Host-buffer path: small calls, chunk schedule, lanes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("small_mode", [1, 2])
@pytest.mark.parametrize("pinned", [True, False])
def test_small_host_calls_complete_by_flag_or_by_stream_sync(ctx, oracle, small_mode, pinned):
    ctx.set_option("small_mode", small_mode)
    flagged0 = ctx.counter("flagged_calls")
    calls = 0
    for n in (1, 2, 255, 256, 1023, 1024, 1025, 4096, 65536, (1 << 18) - 1, 1 << 18):
        words = sxtest.rx_uniform(n, seed=n)
        f = sxtest.tx_gaussian_defined(n, seed=n + 1)
        hw, hf = torch.from_numpy(words), torch.from_numpy(f)
        ho, hi = torch.zeros(2 * n + 2, dtype=torch.float32), torch.zeros(2 * n + 2, dtype=torch.int32)
        if pinned:
            hw, hf, ho, hi = hw.pin_memory(), hf.pin_memory(), ho.pin_memory(), hi.pin_memory()
        for rep in range(2):
            ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
            assert np.array_equal(bits(ho.numpy()[:2 * n]), bits(sxtest.oracle_rx(oracle, words))), (n, rep)
            ctx.convert_tx_buffer_host(hf.data_ptr(), 0, hi.data_ptr(), 0, n, sxtest.THR2_DEFAULT)
            assert np.array_equal(hi.numpy()[:2 * n], sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT)), (n, rep)
            calls += 2
        assert not ho.numpy()[2 * n:].any() and not hi.numpy()[2 * n:].any()
    assert ctx.counter("flagged_calls") - flagged0 == (calls if small_mode == 2 else 0)
    # the CS16 extension takes the same route
    n = 1000
    words = sxtest.rx_uniform(n, seed=77)
    out16 = torch.zeros(2 * n, dtype=torch.int16)
    ctx.convert_rx_buffer_cs16_host(words.ctypes.data, 0, out16.data_ptr(), 0, n)
    assert np.array_equal(out16.numpy(), sxtest.oracle_rx_cs16(oracle, words))


def test_same_staging_buffer_new_contents_every_small_call(ctx, oracle):
    """A driver reuses its pinned staging buffer for every period: no call may see the previous
    call's frames (stale lines in a cache between the GPU and host memory)."""
    n = 256
    hw = torch.zeros(2 * n, dtype=torch.int32).pin_memory()
    ho = torch.zeros(2 * n, dtype=torch.float32).pin_memory()
    for k in range(400):
        words = sxtest.rx_uniform(n, seed=9000 + k)
        hw.copy_(torch.from_numpy(words))
        ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
        assert np.array_equal(bits(ho.numpy()), bits(sxtest.oracle_rx(oracle, words))), k


@pytest.mark.parametrize("c_min,in_mode,out_mode,nt", [(0, 1, 1, 1), (1 << 12, 1, 1, 0), (1 << 16, 1, 1, 1),
                                                       (0, 1, 2, 1), (0, 2, 1, 1), (1 << 12, 2, 2, 0)])
@pytest.mark.parametrize("kinds", ["pinned", "pageable", "pageable_in", "pageable_out"])
@pytest.mark.parametrize("nframes", [(1 << 18) + 1, (1 << 20) + 17, (1 << 23) - 5])
def test_host_pipeline_schedules_and_modes(ctx, oracle, c_min, in_mode, out_mode, nt, kinds, nframes):
    """The copy-engine pipeline with a ramped or uniform chunk schedule, either side moved by the
    copy engine or read/written by the kernel across PCIe, bounce copies with or without
    cache-bypassing stores: same bits whatever the plumbing."""
    ctx.set_option("host_mode", 1)
    ctx.set_option("host_chunk_min_frames", c_min)
    ctx.set_option("host_in_mode", in_mode)
    ctx.set_option("host_out_mode", out_mode)
    ctx.set_option("bounce_nt", nt)
    ctx.set_option("host_chunk_frames", 1 << 19)
    words = sxtest.rx_uniform(nframes, seed=nframes % 1000)
    f = sxtest.tx_uniform(nframes, seed=nframes % 1000 + 1)

    def buf(arr, pageable):
        t = torch.from_numpy(arr)
        return t if pageable else t.pin_memory()

    pin_in = kinds in ("pinned", "pageable_out")
    pin_out = kinds in ("pinned", "pageable_in")
    hw, hf = buf(words, not pin_in), buf(f, not pin_in)
    ho = buf(np.full(2 * nframes + 8, np.float32(7.0)), not pin_out)
    hi = buf(np.full(2 * nframes + 8, 7, np.int32), not pin_out)
    ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, nframes)
    assert np.array_equal(bits(ho.numpy()[:2 * nframes]), bits(sxtest.oracle_rx(oracle, words)))
    ctx.convert_tx_buffer_host(hf.data_ptr(), 0, hi.data_ptr(), 0, nframes, sxtest.THR2_DEFAULT)
    assert np.array_equal(hi.numpy()[:2 * nframes], sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT))
    assert (ho.numpy()[2 * nframes:] == 7.0).all() and (hi.numpy()[2 * nframes:] == 7).all()


@pytest.mark.parametrize("nframes", [256, (1 << 21) + 3])
def test_rx_and_tx_threads_run_side_by_side(ctx, oracle, nframes):
    """One lane per direction (the reference locks per stream, SoapySX.cpp:373, :878, :979): an RX
    thread and a TX thread on one context, pinned and pageable buffers, every result checked."""
    words = sxtest.rx_uniform(nframes, seed=1)
    f = sxtest.tx_uniform(nframes, seed=2)
    want_rx = sxtest.oracle_rx(oracle, words)
    want_tx = sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT)
    reps = 200 if nframes <= 4096 else 6
    errors = []

    def rx_thread(pinned):
        src = torch.from_numpy(words).pin_memory() if pinned else torch.from_numpy(words)
        out = torch.zeros(2 * nframes, dtype=torch.float32)
        out = out.pin_memory() if pinned else out
        for k in range(reps):
            out.zero_()
            ctx.convert_rx_buffer_host(src.data_ptr(), 0, out.data_ptr(), 0, nframes)
            if not np.array_equal(bits(out.numpy()), bits(want_rx)):
                errors.append(("rx", pinned, k))

    def tx_thread(pinned):
        src = torch.from_numpy(f).pin_memory() if pinned else torch.from_numpy(f)
        out = torch.zeros(2 * nframes, dtype=torch.int32)
        out = out.pin_memory() if pinned else out
        for k in range(reps):
            out.zero_()
            ctx.convert_tx_buffer_host(src.data_ptr(), 0, out.data_ptr(), 0, nframes, sxtest.THR2_DEFAULT)
            if not np.array_equal(out.numpy(), want_tx):
                errors.append(("tx", pinned, k))

    for pinned in (True, False):
        ts = [threading.Thread(target=rx_thread, args=(pinned,)), threading.Thread(target=tx_thread, args=(pinned,))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    assert not errors, errors[:5]


def test_resident_converter_leaves_when_the_device_is_synchronised(ctx, oracle):
    """a stream that calls in every period keeps the resident kernel alive; a
    device-wide synchronisation must still return promptly -- at once when the library issues it
    (quit flag), within the kernel's 20 ms lifetime when the application does."""
    from sxxcvr_b200 import Bank
    ctx.set_option("resident_max_frames", 4096)
    n = 256
    words = sxtest.rx_uniform(n, seed=4)
    want = sxtest.oracle_rx(oracle, words)
    hw = torch.from_numpy(words).pin_memory()
    ho = torch.zeros(2 * n, dtype=torch.float32).pin_memory()
    stop = threading.Event()
    bad = []

    def caller():        # a repeater's RX side: one call per 1.7 ms period, for ever
        while not stop.is_set():
            ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
            if not np.array_equal(bits(ho.numpy()), bits(want)):
                bad.append(1)
            time.sleep(0.0017)

    t = threading.Thread(target=caller)
    t.start()
    try:
        time.sleep(0.05)
        worst = 0.0
        for _ in range(10):                      # the application's own device-wide syncs
            t0 = time.perf_counter()
            torch.cuda.synchronize()
            worst = max(worst, time.perf_counter() - t0)
            time.sleep(0.003)
        assert worst < 0.2, worst                # bounded by the 20 ms lifetime (loose: shared box)
        t0 = time.perf_counter()
        with Bank(ctx, 4, 256, 75000.0, 0.0, 1):  # sxgpu_bank_destroy synchronises the device
            pass
        assert time.perf_counter() - t0 < 1.0
    finally:
        stop.set()
        t.join()
    assert not bad
    assert ctx.counter("resident_launches") >= 2


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("nframes", [255, 4096, (1 << 21) + 5])
def test_s16_extension_host_entry_points(ctx, oracle, pinned, nframes):
    """EXTENSION (no reference: SoapySX.cpp:200-207, :474 run 32-bit slots only): the 16-bit I2S
    slot conversions through the host-buffer entry points, against their specification in the oracle."""
    rng = np.random.default_rng(nframes)
    s16 = rng.integers(-32768, 32768, size=2 * nframes, dtype=np.int64).astype(np.int16)
    f = sxtest.tx_gaussian_defined(nframes, seed=nframes)
    hs, hf = torch.from_numpy(s16), torch.from_numpy(f)
    ho = torch.zeros(2 * nframes, dtype=torch.float32)
    hi = torch.zeros(2 * nframes, dtype=torch.int16)
    if pinned:
        hs, hf, ho, hi = hs.pin_memory(), hf.pin_memory(), ho.pin_memory(), hi.pin_memory()
    ctx.convert_rx_buffer_s16_host(hs.data_ptr(), 0, ho.data_ptr(), 0, nframes)
    assert np.array_equal(bits(ho.numpy()), bits(sxtest.oracle_rx_s16(oracle, s16)))
    ctx.convert_tx_buffer_s16_host(hf.data_ptr(), 0, hi.data_ptr(), 0, nframes, sxtest.THR2_DEFAULT)
    assert np.array_equal(hi.numpy(), sxtest.oracle_tx_s16(oracle, f, sxtest.THR2_DEFAULT))
