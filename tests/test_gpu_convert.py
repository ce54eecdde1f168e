"""-m gpu: the CUDA converters, called through the C ABI (include/sxgpu.h), against the CPU
oracle on the same seeded inputs.  Bit-exact everywhere: the tolerance is zero."""
import ctypes as C

import numpy as np
import pytest

import sxtest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def dev(arr: np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def host(t) -> np.ndarray:
    return t.cpu().numpy()


def gpu_rx(ctx, words: np.ndarray, **opts) -> np.ndarray:
    for k, v in opts.items():
        ctx.set_option(k, v)
    src = dev(words.astype(np.int32))
    dst = torch.empty(words.size, dtype=torch.float32, device="cuda")
    ctx.convert_rx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, words.size // 2)
    ctx.stream_sync()
    return host(dst)


def gpu_tx(ctx, floats: np.ndarray, thr2: float, **opts) -> np.ndarray:
    for k, v in opts.items():
        ctx.set_option(k, v)
    src = dev(floats.astype(np.float32))
    dst = torch.empty(floats.size, dtype=torch.int32, device="cuda")
    ctx.convert_tx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, floats.size // 2, thr2)
    ctx.stream_sync()
    return host(dst)


VARIANTS = [dict(rx_variant=4, tx_variant=4), dict(rx_variant=1, tx_variant=1), dict(rx_variant=2, tx_variant=2), dict(rx_variant=3, tx_variant=3),
            dict(rx_variant=3, tx_variant=3, bulk_tile=1024, bulk_stages=4),
            dict(rx_variant=2, tx_variant=2, unroll=8, block=512), dict(rx_variant=1, tx_variant=1, unroll=2)]
RESET = dict(rx_variant=0, tx_variant=0, unroll=0, block=0, bulk_tile=0, bulk_stages=0, ctas_per_sm=0)


@pytest.fixture(autouse=True)
def reset_options(ctx):
    yield
    for k, v in RESET.items():
        ctx.set_option(k, v)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("opts", VARIANTS)
@pytest.mark.parametrize("nframes", [1, 2, 3, 255, 256, 4096, 65536, 65537, 1 << 20, (1 << 20) + 5])
def test_rx_matches_oracle(ctx, oracle, opts, nframes):
    words = sxtest.rx_uniform(nframes, seed=sxtest.SEED + nframes)
    got = gpu_rx(ctx, words, **opts)
    assert np.array_equal(bits(got), bits(sxtest.oracle_rx(oracle, words)))


@pytest.mark.parametrize("opts", VARIANTS[:3])
def test_rx_structured_words(ctx, oracle, opts):
    words = sxtest.rx_structured()
    got = gpu_rx(ctx, words, **opts)
    assert np.array_equal(bits(got), bits(sxtest.oracle_rx(oracle, words)))


def test_rx_survey_kats(ctx):
    # SURVEY.md Appendix A.1, generated from the reference source.
    kat = {0: 0x00000000, 1: 0x30000000, -1: 0xB0000000, 2**31 - 1: 0x3F800000, -2**31: 0xBF800000,
           0x7FFFFF80: 0x3F7FFFFF, 0x7FFFFFBF: 0x3F7FFFFF, 0x7FFFFFC0: 0x3F800000, 16777217: 0x3C000000,
           0x12345678: 0x3E11A2B4, 0x87654321 - 2**32: 0xBF71357A}
    words = np.array(list(kat.keys()) + [0], dtype=np.int64).astype(np.int32)
    got = bits(gpu_rx(ctx, words))
    assert [int(x) for x in got[:len(kat)]] == list(kat.values())


@pytest.mark.parametrize("opts", VARIANTS)
@pytest.mark.parametrize("thr2", [sxtest.THR2_DEFAULT, 0.0])
@pytest.mark.parametrize("nframes", [1, 2, 3, 256, 4097, 65536, (1 << 20) + 3])
def test_tx_uniform_matches_oracle(ctx, oracle, opts, thr2, nframes):
    f = sxtest.tx_uniform(nframes, seed=sxtest.SEED + 1 + nframes)
    got = gpu_tx(ctx, f, thr2, **opts)
    assert np.array_equal(got, sxtest.oracle_tx(oracle, f, thr2))


@pytest.mark.parametrize("opts", VARIANTS[:3])
def test_tx_gaussian_with_negative_clamp(ctx, oracle, opts):
    f = sxtest.tx_gaussian_defined(1 << 18)
    assert (f <= -1.0).sum() > 1000           # the clamp is exercised
    got = gpu_tx(ctx, f, sxtest.THR2_DEFAULT, **opts)
    assert np.array_equal(got, sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT))


@pytest.mark.parametrize("opts", VARIANTS[:3])
@pytest.mark.parametrize("thr2", [sxtest.THR2_DEFAULT, 0.25, 1.0e-6])
def test_tx_threshold_circle_is_unfused(ctx, oracle, opts, thr2):
    f = sxtest.tx_threshold_circle(1 << 18, thr2)
    want = sxtest.oracle_tx(oracle, f, thr2)
    flags = (want[0::2] & 3)
    assert 0.2 < (flags == 3).mean() < 0.8    # the set really straddles the circle
    got = gpu_tx(ctx, f, thr2, **opts)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("opts", VARIANTS[:3])
@pytest.mark.parametrize("thr2", [sxtest.THR2_DEFAULT, 0.0, float("nan"), float("inf")])
def test_tx_undefined_domain_follows_arm_semantics(ctx, oracle, opts, thr2):
    f = sxtest.tx_specials()
    got = gpu_tx(ctx, f, thr2, **opts)
    assert np.array_equal(got, sxtest.oracle_tx(oracle, f, thr2))


def test_tx_survey_kats(ctx):
    # SURVEY.md Appendix A.2 (x86 reference build on the defined domain; ARM answers beyond it).
    thr = sxtest.THR2_DEFAULT
    cases = [((-1.0, -1.0), (0x80000003, 0x80000000)),
             ((0.99999994, -0.99999994), (0x7FFFFF83, 0x80000080)),
             ((0.3, -0.3), (0x26666683, 0xD9999980)),
             ((1e-3, 0.0), (0x0020C49B, 0)),
             ((0.000707106781, 0.000707106781), (0x00172BA4, 0x00172BA4)),
             ((0.0, -0.0), (0, 0)),
             ((1.0, 0.0), (0x7FFFFFFF, 0)),          # test_timestamps.py:34 burst, ARM saturation
             ((2.0, float("inf")), (0x7FFFFFFF, 0x7FFFFFFC)),
             ((float("nan"), 0.5), (0, 0x40000000))]
    f = np.array([c[0] for c in cases], dtype=np.float32).ravel()
    got = bits(gpu_tx(ctx, f, thr)).reshape(-1, 2)
    for (inp, want), g in zip(cases, got):
        assert (int(g[0]), int(g[1])) == want, inp
    z = bits(gpu_tx(ctx, np.array([0.0, -0.0, 0.000707106781, 0.000707106781], np.float32), 0.0))
    assert [int(x) for x in z] == [3, 0, 0x00172BA7, 0x00172BA4]


@pytest.mark.parametrize("src_off,dst_off", [(0, 0), (1, 1), (1, 0), (0, 1), (3, 3), (3, 2), (2, 1), (5, 7)])
@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_frame_offsets_and_misaligned_views(ctx, oracle, src_off, dst_off, variant):
    """Offsets are in frames, as in the reference (SoapySX.cpp:105-106, :118-119); an odd frame
    offset leaves only 8-byte alignment and different src/dest offsets defeat wide vectors."""
    n = 10007
    words = sxtest.rx_uniform(n + 8)
    ctx.set_option("rx_variant", variant)
    ctx.set_option("tx_variant", variant)
    src = dev(words)
    dst = torch.full((2 * (n + 8),), -7.0, dtype=torch.float32, device="cuda")
    ctx.convert_rx_buffer(src.data_ptr(), src_off, dst.data_ptr(), dst_off, n)
    ctx.stream_sync()
    got = host(dst)
    want = sxtest.oracle_rx(oracle, words[2 * src_off: 2 * (src_off + n)])
    assert np.array_equal(bits(got[2 * dst_off: 2 * (dst_off + n)]), bits(want))
    assert (got[:2 * dst_off] == -7.0).all() and (got[2 * (dst_off + n):] == -7.0).all()   # no overrun

    f = sxtest.tx_uniform(n + 8)
    fsrc = dev(f)
    idst = torch.full((2 * (n + 8),), 0x55, dtype=torch.int32, device="cuda")
    ctx.convert_tx_buffer(fsrc.data_ptr(), src_off, idst.data_ptr(), dst_off, n, sxtest.THR2_DEFAULT)
    ctx.stream_sync()
    goti = host(idst)
    wanti = sxtest.oracle_tx(oracle, f[2 * src_off: 2 * (src_off + n)], sxtest.THR2_DEFAULT)
    assert np.array_equal(goti[2 * dst_off: 2 * (dst_off + n)], wanti)
    assert (goti[:2 * dst_off] == 0x55).all() and (goti[2 * (dst_off + n):] == 0x55).all()


def test_word_aligned_pointers_take_the_word_kernel(ctx, oracle):
    n = 4099
    words = sxtest.rx_uniform(n + 2)
    src = dev(words)
    dst = torch.zeros(2 * (n + 2), dtype=torch.float32, device="cuda")
    ctx.convert_rx_buffer(src.data_ptr() + 4, 0, dst.data_ptr() + 4, 0, n)     # 4-byte aligned only
    ctx.stream_sync()
    assert np.array_equal(bits(host(dst)[1:1 + 2 * n]), bits(sxtest.oracle_rx(oracle, words[1:1 + 2 * n])))


def test_in_place_conversion(ctx, oracle):
    n = 1 << 16
    words = sxtest.rx_uniform(n)
    buf = dev(words)
    ctx.convert_rx_buffer(buf.data_ptr(), 0, buf.data_ptr(), 0, n)
    ctx.stream_sync()
    assert np.array_equal(bits(host(buf)), bits(sxtest.oracle_rx(oracle, words)))


def test_zero_length_and_bad_arguments(ctx):
    from sxxcvr_b200 import SxGpuError
    ctx.convert_rx_buffer(0, 0, 0, 0, 0)         # length 0 is legal (SoapySX.cpp:1090 does it)
    ctx.convert_tx_buffer(0, 0, 0, 0, 0, 0.0)
    t = torch.zeros(16, dtype=torch.int32, device="cuda")
    with pytest.raises(SxGpuError):
        ctx.convert_rx_buffer(0, 0, t.data_ptr(), 0, 4)
    with pytest.raises(SxGpuError):
        ctx.convert_rx_buffer(t.data_ptr() + 2, 0, t.data_ptr(), 0, 4)
    with pytest.raises(SxGpuError):
        ctx.convert_rx_buffer(t.data_ptr(), 1 << 61, t.data_ptr(), 0, 4)
    with pytest.raises(SxGpuError):
        ctx.set_option("no_such_option", 1)


def test_loopback_is_tx_of_rx_not_identity(ctx, oracle):
    n = (1 << 18) + 1
    words = sxtest.rx_uniform(n)
    src = dev(words)
    mid = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    out = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    ctx.convert_loopback(src.data_ptr(), mid.data_ptr(), out.data_ptr(), n, 0.0)
    ctx.stream_sync()
    want_mid = sxtest.oracle_rx(oracle, words)
    want_out = sxtest.oracle_tx(oracle, want_mid, 0.0)
    assert np.array_equal(bits(host(mid)), bits(want_mid))
    assert np.array_equal(host(out), want_out)
    assert (host(out) != words).mean() > 0.5      # float keeps 24 bits
    out2 = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    ctx.convert_loopback(src.data_ptr(), None, out2.data_ptr(), n, 0.0)
    ctx.stream_sync()
    assert np.array_equal(host(out2), want_out)


@pytest.mark.parametrize("nblocks,length", [(1, 256), (64, 256), (4096, 256), (1000, 255), (37, 4096), (5, 70001)])
def test_batched_blocks(ctx, oracle, nblocks, length):
    stride = length + 3          # odd strides give 8-byte-aligned blocks too
    words = sxtest.rx_uniform(nblocks * stride)
    src = dev(words)
    dst = torch.zeros(2 * nblocks * stride, dtype=torch.float32, device="cuda")
    from sxxcvr_b200.capi import Block
    blocks = [Block(src.data_ptr() + 8 * b * stride, dst.data_ptr() + 8 * b * stride, length, 0.0, 0)
              for b in range(nblocks)]
    ctx.convert_batch("rx", blocks)
    ctx.stream_sync()
    got = host(dst).reshape(nblocks, 2 * stride)
    w = words.reshape(nblocks, 2 * stride)
    for b in (0, nblocks // 2, nblocks - 1):
        assert np.array_equal(bits(got[b, :2 * length]), bits(sxtest.oracle_rx(oracle, w[b, :2 * length])))
        assert (got[b, 2 * length:] == 0).all()
    want_all = sxtest.oracle_rx(oracle, words).reshape(nblocks, 2 * stride)[:, :2 * length]
    assert np.array_equal(bits(got[:, :2 * length]), bits(want_all))

    f = sxtest.tx_uniform(nblocks * stride)
    fsrc = dev(f)
    idst = torch.zeros(2 * nblocks * stride, dtype=torch.int32, device="cuda")
    thr = [sxtest.THR2_DEFAULT if b % 2 else 0.0 for b in range(nblocks)]
    tblocks = [Block(fsrc.data_ptr() + 8 * b * stride, idst.data_ptr() + 8 * b * stride, length, thr[b], 0)
               for b in range(nblocks)]
    ctx.convert_batch("tx", tblocks)
    ctx.stream_sync()
    goti = host(idst).reshape(nblocks, 2 * stride)
    fr = f.reshape(nblocks, 2 * stride)
    for b in range(0, nblocks, max(1, nblocks // 16)):
        assert np.array_equal(goti[b, :2 * length], sxtest.oracle_tx(oracle, fr[b, :2 * length], thr[b])), b


def test_cs16_extensions_match_their_specification(ctx, oracle):
    """EXTENSION: no reference implementation exists (SoapySX.cpp:752-753); parity is against
    our own scalar specification in oracle/sx_oracle.c."""
    for n in (1, 2, 3, 4, 5, 4096, 100003):
        for variant in (1, 2, 3, 4):
            ctx.set_option("rx_variant", variant)
            ctx.set_option("tx_variant", variant)
            words = sxtest.rx_uniform(n, seed=n)
            src = dev(words)
            dst = torch.zeros(2 * n, dtype=torch.int16, device="cuda")
            ctx.convert_rx_buffer_cs16(src.data_ptr(), 0, dst.data_ptr(), 0, n)
            ctx.stream_sync()
            assert np.array_equal(host(dst), sxtest.oracle_rx_cs16(oracle, words)), (n, variant)

            rng = np.random.default_rng(n)
            s = rng.integers(-32768, 32768, size=2 * n, dtype=np.int64).astype(np.int16)
            ssrc = dev(s)
            idst = torch.zeros(2 * n, dtype=torch.int32, device="cuda")
            ctx.convert_tx_buffer_cs16(ssrc.data_ptr(), 0, idst.data_ptr(), 0, n, sxtest.THR2_DEFAULT)
            ctx.stream_sync()
            assert np.array_equal(host(idst), sxtest.oracle_tx_cs16(oracle, s, sxtest.THR2_DEFAULT)), (n, variant)


def test_s16_frame_extension_matches_its_specification(ctx, oracle):
    """EXTENSION: 16-bit I2S slots; no reference implementation (SoapySX.cpp:200-207, :474)."""
    for n in (1, 2, 3, 4, 7, 4096, 100003, (1 << 21) + 2):
        for variant in (1, 2, 3, 4):
            ctx.set_option("rx_variant", variant)
            ctx.set_option("tx_variant", variant)
            rng = np.random.default_rng(n)
            s = rng.integers(-32768, 32768, size=2 * n, dtype=np.int64).astype(np.int16)
            src = dev(s)
            dst = torch.zeros(2 * n, dtype=torch.float32, device="cuda")
            ctx.convert_rx_buffer_s16(src.data_ptr(), 0, dst.data_ptr(), 0, n)
            ctx.stream_sync()
            assert np.array_equal(bits(host(dst)), bits(sxtest.oracle_rx_s16(oracle, s))), (n, variant)
            f = np.concatenate([sxtest.tx_gaussian_defined(n, seed=n), sxtest.tx_specials()])[: 2 * n]
            if f.size < 2 * n:
                f = np.resize(f, 2 * n)
            fsrc = dev(f)
            out = torch.zeros(2 * n, dtype=torch.int16, device="cuda")
            for thr2 in (sxtest.THR2_DEFAULT, 0.0):
                ctx.convert_tx_buffer_s16(fsrc.data_ptr(), 0, out.data_ptr(), 0, n, thr2)
                ctx.stream_sync()
                assert np.array_equal(host(out), sxtest.oracle_tx_s16(oracle, f, thr2)), (n, variant, thr2)


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("nframes", [255, 4096, (1 << 21) + 5])
def test_cs16_extension_host_entry_points(ctx, oracle, pinned, nframes):
    words = sxtest.rx_uniform(nframes, seed=9)
    hw = torch.from_numpy(words)
    ho = torch.zeros(2 * nframes, dtype=torch.int16)
    hi = torch.zeros(2 * nframes, dtype=torch.int32)
    if pinned:
        hw, ho, hi = hw.pin_memory(), ho.pin_memory(), hi.pin_memory()
    ctx.convert_rx_buffer_cs16_host(hw.data_ptr(), 0, ho.data_ptr(), 0, nframes)
    want = sxtest.oracle_rx_cs16(oracle, words)
    assert np.array_equal(ho.numpy(), want)
    ctx.convert_tx_buffer_cs16_host(ho.data_ptr(), 0, hi.data_ptr(), 0, nframes, sxtest.THR2_DEFAULT)
    assert np.array_equal(hi.numpy(), sxtest.oracle_tx_cs16(oracle, want, sxtest.THR2_DEFAULT))


def test_synth_frames_and_stats_match_the_host_definitions(ctx, oracle):
    n = 100001
    buf = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    ctx.synth_frames(buf.data_ptr(), 12345, n, sxtest.SEED)
    ctx.stream_sync()
    want = sxtest.synth_frames(oracle, 12345, n)
    assert np.array_equal(host(buf), want)
    assert ctx.stats_words(buf.data_ptr(), 2 * n, 77) == sxtest.oracle_stats(oracle, want, 77)
    ctx.fill_silence(buf.data_ptr(), 5, n - 10)
    ctx.stream_sync()
    got = host(buf)
    assert (got[10:2 * (n - 5)] == 0).all() and np.array_equal(got[:10], want[:10]) and np.array_equal(got[-10:], want[-10:])


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("mode,nframes", [(0, 256), (2, 4096), (1, 256), (1, (1 << 22) + 17), (0, (1 << 23) + 1)])
def test_host_buffer_entry_points(ctx, oracle, pinned, mode, nframes):
    """sxgpu_convert_*_buffer_host: the call shape of the reference's own call sites
    (host staging vector <-> caller's host buffer, SoapySX.cpp:957, :1090)."""
    ctx.set_option("host_mode", mode)
    ctx.set_option("host_chunk_frames", 1 << 20)
    try:
        words = sxtest.rx_uniform(nframes, seed=nframes)
        f = sxtest.tx_uniform(nframes, seed=nframes + 1)
        if pinned:
            hw = torch.from_numpy(words).pin_memory()
            hf = torch.from_numpy(f).pin_memory()
            ho = torch.empty(2 * nframes, dtype=torch.float32).pin_memory()
            hi = torch.empty(2 * nframes, dtype=torch.int32).pin_memory()
        else:
            hw, hf = torch.from_numpy(words), torch.from_numpy(f)
            ho = torch.empty(2 * nframes, dtype=torch.float32)
            hi = torch.empty(2 * nframes, dtype=torch.int32)
        ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, nframes)
        assert np.array_equal(bits(ho.numpy()), bits(sxtest.oracle_rx(oracle, words)))
        ctx.convert_tx_buffer_host(hf.data_ptr(), 0, hi.data_ptr(), 0, nframes, sxtest.THR2_DEFAULT)
        assert np.array_equal(hi.numpy(), sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT))
    finally:
        ctx.set_option("host_mode", 0)
        ctx.set_option("host_chunk_frames", 0)


@pytest.mark.parametrize("threads", [1, 2, 7])
def test_pageable_callers_with_shared_bounce_copies(ctx, oracle, threads):
    """Pageable caller buffers (numpy arrays: what the reference's Python callers pass) are moved
    through pinned staging by `bounce_threads` threads; the result does not depend on how many."""
    nframes = (1 << 22) + 17
    ctx.set_option("bounce_threads", threads)
    ctx.set_option("host_chunk_frames", (1 << 20) + 5)      # chunks that do not split evenly
    try:
        words = sxtest.rx_uniform(nframes, seed=21)
        f = sxtest.tx_uniform(nframes, seed=22)
        out_f = np.full(2 * nframes + 8, np.float32(7.0))   # guard elements behind the block
        out_i = np.full(2 * nframes + 8, 7, np.int32)
        ctx.convert_rx_buffer_host(words.ctypes.data, 0, out_f.ctypes.data, 0, nframes)
        assert np.array_equal(bits(out_f[:2 * nframes]), bits(sxtest.oracle_rx(oracle, words)))
        ctx.convert_tx_buffer_host(f.ctypes.data, 0, out_i.ctypes.data, 0, nframes, sxtest.THR2_DEFAULT)
        assert np.array_equal(out_i[:2 * nframes], sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT))
        assert (out_f[2 * nframes:] == 7.0).all() and (out_i[2 * nframes:] == 7).all()
    finally:
        ctx.set_option("bounce_threads", 0)
        ctx.set_option("host_chunk_frames", 0)


@pytest.mark.parametrize("nframes", [256, (1 << 19) + 3, (1 << 22) + 1])
def test_host_entry_points_accept_device_memory_on_either_side(ctx, oracle, nframes):
    """A torch/cupy buffer handed to readStream/writeStream: that side's PCIe copy is skipped."""
    words = sxtest.rx_uniform(nframes, seed=3)
    want = sxtest.oracle_rx(oracle, words)
    host_in = torch.from_numpy(words).pin_memory()
    dev_in = host_in.cuda()
    dev_out = torch.zeros(2 * nframes, dtype=torch.float32, device="cuda")
    host_out = torch.zeros(2 * nframes, dtype=torch.float32).pin_memory()
    pageable_out = torch.zeros(2 * nframes, dtype=torch.float32)
    h2d0, d2h0 = ctx.counter("h2d_bytes"), ctx.counter("d2h_bytes")
    ctx.convert_rx_buffer_host(host_in.data_ptr(), 0, dev_out.data_ptr(), 0, nframes)      # host -> device
    assert np.array_equal(bits(host(dev_out)), bits(want))
    assert (ctx.counter("h2d_bytes") - h2d0, ctx.counter("d2h_bytes") - d2h0) == (8 * nframes, 0)
    ctx.convert_rx_buffer_host(dev_in.data_ptr(), 0, host_out.data_ptr(), 0, nframes)      # device -> pinned host
    assert np.array_equal(bits(host_out.numpy()), bits(want))
    ctx.convert_rx_buffer_host(dev_in.data_ptr(), 0, pageable_out.data_ptr(), 0, nframes)  # device -> pageable host
    assert np.array_equal(bits(pageable_out.numpy()), bits(want))
    dev_out.zero_()
    ctx.convert_rx_buffer_host(dev_in.data_ptr(), 0, dev_out.data_ptr(), 0, nframes)       # device -> device
    assert np.array_equal(bits(host(dev_out)), bits(want))
    f = sxtest.tx_uniform(nframes, seed=4)
    dev_f = torch.from_numpy(f).cuda()
    out_i = torch.zeros(2 * nframes, dtype=torch.int32).pin_memory()
    ctx.convert_tx_buffer_host(dev_f.data_ptr(), 0, out_i.data_ptr(), 0, nframes, sxtest.THR2_DEFAULT)
    assert np.array_equal(out_i.numpy(), sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT))


def test_resident_converter_for_period_sized_blocks(ctx, oracle):
    """Opt-in low-latency path: a resident kernel rung through a doorbell in pinned memory.  Same
    bits as every other path; survives its own idle timeout (it leaves after 2 ms and is relaunched);
    large blocks still take the normal route."""
    import time
    ctx.set_option("resident_max_frames", 4096)
    try:
        calls0, launches0 = ctx.counter("resident_calls"), ctx.counter("resident_launches")
        for pinned in (True, False):
            for n in (1, 2, 255, 256, 1000, 4096):
                words = sxtest.rx_uniform(n, seed=n)
                f = sxtest.tx_gaussian_defined(n, seed=n + 1)
                hw, hf = torch.from_numpy(words), torch.from_numpy(f)
                ho, hi = torch.zeros(2 * n, dtype=torch.float32), torch.zeros(2 * n, dtype=torch.int32)
                if pinned:
                    hw, hf, ho, hi = hw.pin_memory(), hf.pin_memory(), ho.pin_memory(), hi.pin_memory()
                for rep in range(3):
                    ho.zero_()
                    ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
                    assert np.array_equal(bits(ho.numpy()), bits(sxtest.oracle_rx(oracle, words))), (pinned, n, rep)
                    ctx.convert_tx_buffer_host(hf.data_ptr(), 0, hi.data_ptr(), 0, n, sxtest.THR2_DEFAULT)
                    assert np.array_equal(hi.numpy(), sxtest.oracle_tx(oracle, f, sxtest.THR2_DEFAULT)), (pinned, n, rep)
        assert ctx.counter("resident_calls") - calls0 == 2 * 6 * 3 * 2
        # the same staging buffer reused with new contents every call: no stale cache lines
        n = 256
        hw = torch.zeros(2 * n, dtype=torch.int32).pin_memory()
        ho = torch.zeros(2 * n, dtype=torch.float32).pin_memory()
        for k in range(500):
            words = sxtest.rx_uniform(n, seed=5000 + k)
            hw.copy_(torch.from_numpy(words))
            ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
            assert np.array_equal(bits(ho.numpy()), bits(sxtest.oracle_rx(oracle, words))), k
        # idle longer than the kernel's timeout: it leaves, the next call starts a new one
        before = ctx.counter("resident_launches")
        for _ in range(3):
            time.sleep(0.02)
            ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
            assert np.array_equal(bits(ho.numpy()), bits(sxtest.oracle_rx(oracle, words)))
        assert ctx.counter("resident_launches") - before >= 2
        # bigger than the limit: the usual path, interleaved with resident calls
        big = sxtest.rx_uniform(100000, seed=1)
        hb = torch.from_numpy(big).pin_memory()
        hob = torch.zeros(200000, dtype=torch.float32).pin_memory()
        c = ctx.counter("resident_calls")
        ctx.convert_rx_buffer_host(hb.data_ptr(), 0, hob.data_ptr(), 0, 100000)
        ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
        assert ctx.counter("resident_calls") - c == 1
        assert np.array_equal(bits(hob.numpy()), bits(sxtest.oracle_rx(oracle, big)))
        # latency, for the record (asserted loosely: it must not be slower than the launch path)
        t0 = time.perf_counter()
        for _ in range(2000):
            ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
        resident_us = (time.perf_counter() - t0) / 2000 * 1e6
        ctx.set_option("resident_max_frames", 0)
        t0 = time.perf_counter()
        for _ in range(2000):
            ctx.convert_rx_buffer_host(hw.data_ptr(), 0, ho.data_ptr(), 0, n)
        launch_us = (time.perf_counter() - t0) / 2000 * 1e6
        print(f"256-frame RX call: resident {resident_us:.2f} us, launch+sync {launch_us:.2f} us")
        assert resident_us < launch_us * 3      # sanity only: timing on a shared box is noisy
    finally:
        ctx.set_option("resident_max_frames", 0)


def test_full_size_block_by_properties(ctx, oracle):
    """BASELINE config 5 size (2^27 frames = 1 GiB in): too big for the scalar oracle in a test,
    so check size-independent properties: the checksum of the output equals the checksum of
    chunk-wise oracle output on sampled chunks, RX is odd (rx(-x) == -rx(x)), and TX(RX(x)) keeps
    the top 24 bits' worth of x."""
    n = 1 << 27
    src = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    ctx.synth_frames(src.data_ptr(), 0, n, sxtest.SEED)
    dst = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    ctx.convert_rx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, n)
    ctx.stream_sync()
    # sampled chunks against the oracle, including the very end
    for first in (0, 12345678, n - 4096):
        want = sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, first, 4096))
        assert np.array_equal(bits(host(dst[2 * first: 2 * (first + 4096)])), bits(want))
    # whole-buffer checksum equals the sum of per-shard checksums (shards reduce in any order)
    whole = ctx.stats_words(dst.data_ptr(), 2 * n, 0)
    half = n   # words
    a = ctx.stats_words(dst.data_ptr(), half, 0)
    b = ctx.stats_words(dst.data_ptr() + 4 * half, 2 * n - half, half)
    M = (1 << 64) - 1
    assert whole == ((a[0] + b[0]) & M, (a[1] + b[1]) & M, a[2] ^ b[2], a[3] + b[3], a[4] + b[4], a[5] + b[5])
    # magnitude bound and linearity in the exponent: |rx| <= 1 everywhere
    assert float(dst.abs().max()) <= 1.0
    # TX(RX(x)) differs from x by less than 2^8 + 4 in every word (24-bit significand, 2 flag bits)
    out = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    ctx.convert_tx_buffer(dst.data_ptr(), 0, out.data_ptr(), 0, n, 2.0)
    ctx.stream_sync()
    diff = (out.to(torch.int64) - src.to(torch.int64)).abs()
    # the rail: RX(INT32_MAX-ish) = 1.0 -> saturates to 0x7FFFFFFC
    assert int(diff.max()) <= 260
    del diff
    assert ctx.stats_words(out.data_ptr(), 2 * n, 0)[4] == 0      # thr2 = 2.0 > max |z|^2: PA never enabled


def test_maximum_block_of_the_sweep_4gib(ctx, oracle):
    """BASELINE config 5's largest block: 2^29 frames = 4 GiB on the input side, 2^30 words, byte
    offsets beyond 32 bits.  Sampled chunks (start, the 4 GiB-byte boundary, an odd place, the very
    end) against the oracle in both directions, and shard-wise checksums that must add up."""
    n = 1 << 29
    i2s = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    cf = torch.empty(2 * n, dtype=torch.float32, device="cuda")
    ctx.synth_frames(i2s.data_ptr(), 0, n, 99)
    ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n)
    ctx.stream_sync()
    spots = (0, (1 << 28) - 2048, (1 << 28), 333333333, n - 4096)
    for first in spots:
        frames = sxtest.synth_frames(oracle, first, 4096, seed=99)
        assert np.array_equal(host(i2s[2 * first: 2 * (first + 4096)]), frames), first
        want = sxtest.oracle_rx(oracle, frames)
        assert np.array_equal(bits(host(cf[2 * first: 2 * (first + 4096)])), bits(want)), first
    out = i2s          # reuse the input buffer for the TX result
    ctx.convert_tx_buffer(cf.data_ptr(), 0, out.data_ptr(), 0, n, sxtest.THR2_DEFAULT)
    ctx.stream_sync()
    for first in spots:
        want_cf = sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, first, 4096, seed=99))
        assert np.array_equal(host(out[2 * first: 2 * (first + 4096)]), sxtest.oracle_tx(oracle, want_cf, sxtest.THR2_DEFAULT)), first
    whole = ctx.stats_words(out.data_ptr(), 2 * n, 0)
    cut = (1 << 29) + 6        # words; an unaligned cut
    a = ctx.stats_words(out.data_ptr(), cut, 0)
    b = ctx.stats_words(out.data_ptr() + 4 * cut, 2 * n - cut, cut)
    from sxxcvr_b200 import sharding
    assert whole == sharding.combine_stats([a, b]) and whole[3] == 2 * n
