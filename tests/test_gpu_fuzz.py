"""-m gpu: randomised shapes.  Lengths, frame offsets, byte misalignments, schedules and options
drawn at random (fixed seed); every call is compared with the oracle and the bytes around the
destination must be untouched."""
import numpy as np
import pytest

import sxtest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

OPTIONS = dict(rx_variant=(0, 1, 2, 3, 4), tx_variant=(0, 1, 2, 3, 4), unroll=(0, 2, 4, 8), block=(0, 256, 512),
               ctas_per_sm=(0, 1, 3), bulk_tile=(0, 512, 1024, 2048), bulk_stages=(0, 4))


def draw_length(rng):
    kind = rng.integers(0, 4)
    if kind == 0:
        return int(rng.integers(1, 64))
    if kind == 1:
        return int(rng.integers(64, 5000))
    if kind == 2:
        return int(2 ** rng.integers(8, 18) + rng.integers(-3, 4))
    return int(rng.integers(100000, 700000))


def test_random_shapes_offsets_and_schedules(ctx, oracle):
    rng = np.random.default_rng(20261017)
    cap = 700000 + 64
    words = sxtest.rx_uniform(cap, seed=11)
    floats = sxtest.tx_gaussian_defined(cap, seed=12)
    d_words = torch.from_numpy(words).cuda()
    d_floats = torch.from_numpy(floats).cuda()
    try:
        import os
        for case in range(int(os.environ.get("SX_FUZZ_CASES", "300"))):
            opts = {k: int(rng.choice(v)) for k, v in OPTIONS.items()}
            if opts["bulk_tile"] == 512:
                opts["bulk_stages"] = 4
            if opts["bulk_tile"] == 0:
                opts["bulk_stages"] = 0
            for k, v in opts.items():
                ctx.set_option(k, v)
            n = draw_length(rng)
            so, do = int(rng.integers(0, 9)), int(rng.integers(0, 9))
            sb, db = int(rng.choice((0, 0, 0, 4))), int(rng.choice((0, 0, 0, 4)))      # byte misalignment
            thr2 = float(rng.choice((0.0, sxtest.THR2_DEFAULT, 0.25, 1.5)))
            tag = (case, opts, n, so, do, sb, db, thr2)

            out = torch.full((2 * (n + do) + 8,), 7.0, dtype=torch.float32, device="cuda")
            ctx.convert_rx_buffer(d_words.data_ptr() + sb, so, out.data_ptr() + db, do, n)
            ctx.stream_sync()
            got = out.cpu().numpy()
            w0 = 2 * so + sb // 4
            want = sxtest.oracle_rx(oracle, words[w0: w0 + 2 * n])
            o0 = 2 * do + db // 4
            assert np.array_equal(got[o0: o0 + 2 * n].view(np.uint32), want.view(np.uint32)), tag
            assert (got[:o0] == 7.0).all() and (got[o0 + 2 * n:] == 7.0).all(), tag

            outi = torch.full((2 * (n + do) + 8,), 0x5A5A, dtype=torch.int32, device="cuda")
            ctx.convert_tx_buffer(d_floats.data_ptr() + sb, so, outi.data_ptr() + db, do, n, thr2)
            ctx.stream_sync()
            goti = outi.cpu().numpy()
            wanti = sxtest.oracle_tx(oracle, floats[w0: w0 + 2 * n], thr2)
            assert np.array_equal(goti[o0: o0 + 2 * n], wanti), tag
            assert (goti[:o0] == 0x5A5A).all() and (goti[o0 + 2 * n:] == 0x5A5A).all(), tag
    finally:
        for k in OPTIONS:
            ctx.set_option(k, 0)


def test_concurrent_host_threads_on_their_own_streams(ctx, oracle):
    """The C ABI is documented as callable from several host threads, each with its own stream."""
    import threading
    from sxxcvr_b200 import capi
    import ctypes as C
    n = 1 << 16
    errors = []

    def worker(tid):
        try:
            st = C.c_void_p()
            ctx.check(ctx.lib.sxgpu_stream_create(ctx.handle, C.byref(st)), "stream_create")
            words = sxtest.rx_uniform(n, seed=100 + tid)
            want = sxtest.oracle_rx(oracle, words)
            want_tx = sxtest.oracle_tx(oracle, want, 0.0)
            src = torch.from_numpy(words).cuda()
            mid = torch.empty(2 * n, dtype=torch.float32, device="cuda")
            dst = torch.empty(2 * n, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            for _ in range(50):
                ctx.convert_rx_buffer(src.data_ptr(), 0, mid.data_ptr(), 0, n, st)
                ctx.convert_tx_buffer(mid.data_ptr(), 0, dst.data_ptr(), 0, n, 0.0, st)
            ctx.stream_sync(st)
            if not np.array_equal(mid.cpu().numpy().view(np.uint32), want.view(np.uint32)):
                errors.append((tid, "rx"))
            if not np.array_equal(dst.cpu().numpy(), want_tx):
                errors.append((tid, "tx"))
            ctx.lib.sxgpu_stream_destroy(ctx.handle, st)
        except Exception as e:       # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
