"""-m gpu: exhaustive parity.  Every one of the 2^32 possible I2S words through RX, and every one
of the 2^32 float bit patterns through each TX component (NaNs, infinities, subnormals, both
signs), CUDA versus the C oracle, bit for bit.  Chunked: 2^27 values per chunk on the GPU, the
oracle spread over the host cores (ctypes releases the GIL)."""
import os
import threading

import numpy as np
import pytest

import sxtest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CHUNK = 1 << 27          # values per chunk
NCHUNKS = (1 << 32) // CHUNK


def parallel(fn, total, nthreads):
    per = -(-total // nthreads)
    per += per & 1
    ts = [threading.Thread(target=fn, args=(a, min(per, total - a))) for a in range(0, total, per)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


@pytest.fixture(scope="module")
def nthreads():
    return max(1, len(os.sched_getaffinity(0)))


def test_rx_every_possible_word(ctx, oracle, nthreads):
    words_dev = torch.empty(CHUNK, dtype=torch.int32, device="cuda")
    out_dev = torch.empty(CHUNK, dtype=torch.float32, device="cuda")
    want = np.empty(CHUNK, np.float32)
    base = torch.arange(CHUNK, dtype=torch.int64, device="cuda")
    for c in range(NCHUNKS):
        # chunk c holds the words c*CHUNK .. (c+1)*CHUNK-1 taken as unsigned 32-bit patterns
        words_dev.copy_(((base + c * CHUNK + 2**31) % 2**32 - 2**31).to(torch.int32))
        torch.cuda.synchronize()      # the context's stream does not order after torch's default stream
        ctx.convert_rx_buffer(words_dev.data_ptr(), 0, out_dev.data_ptr(), 0, CHUNK // 2)
        ctx.stream_sync()
        words = words_dev.cpu().numpy()
        got = out_dev.cpu().numpy()

        def work(first, n):
            oracle.sxo_convert_rx_buffer(words.ctypes.data, first // 2, want.ctypes.data, first // 2, n // 2)

        parallel(work, CHUNK, nthreads)
        if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
            bad = int(np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))[0])
            pytest.fail(f"RX differs for word {int(words[bad]):#x}: {got.view(np.uint32)[bad]:#x} != {want.view(np.uint32)[bad]:#x}")


@pytest.mark.parametrize("slot,other,thr2", [("I", 0.25, sxtest.THR2_DEFAULT), ("Q", -0.75, 0.5625)])
def test_tx_every_possible_float(ctx, oracle, nthreads, slot, other, thr2):
    """The swept component takes every float bit pattern; the other is fixed so that the threshold
    comparison is exercised on both sides (|other|^2 is below thr2 in the first case and exactly
    equal to it in the second)."""
    nframes = CHUNK
    f_dev = torch.empty(2 * nframes, dtype=torch.float32, device="cuda")
    out_dev = torch.empty(2 * nframes, dtype=torch.int32, device="cuda")
    want = np.empty(2 * nframes, np.int32)
    base = torch.arange(nframes, dtype=torch.int64, device="cuda")
    pairs = f_dev.view(nframes, 2)
    sweep_col, fixed_col = (0, 1) if slot == "I" else (1, 0)
    pairs[:, fixed_col] = other
    for c in range(NCHUNKS):
        bits = ((base + c * CHUNK + 2**31) % 2**32 - 2**31).to(torch.int32)
        pairs[:, sweep_col] = bits.view(torch.float32)
        torch.cuda.synchronize()
        ctx.convert_tx_buffer(f_dev.data_ptr(), 0, out_dev.data_ptr(), 0, nframes, thr2)
        ctx.stream_sync()
        f = f_dev.cpu().numpy()
        got = out_dev.cpu().numpy()

        def work(first, n):
            oracle.sxo_convert_tx_buffer(f.ctypes.data, first, want.ctypes.data, first, n, thr2)

        parallel(work, nframes, nthreads)
        if not np.array_equal(got, want):
            bad = int(np.flatnonzero(got != want)[0])
            pytest.fail(f"TX {slot} differs for input bits {f.view(np.uint32)[bad]:#x}: "
                        f"{got.view(np.uint32)[bad]:#x} != {want.view(np.uint32)[bad]:#x}")
