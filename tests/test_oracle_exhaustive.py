"""CPU: the C oracle against the reference's own converters (oracle/_ref) over EVERY input.

RX: all 2^32 I2S words.  TX: all 2^32 float bit patterns in the I slot (Q fixed) and in the Q slot
(I fixed), compared wherever the reference is defined C++ (the swept component not NaN and < 1.0);
outside that domain the x86 reference build is not an arbiter (parity policy) and the
oracle is checked against the independent numpy restatement instead (test_oracle.py).
Chunked and spread over the host cores (ctypes releases the GIL).

About 8 minutes on 8 cores, so it only runs when SX_EXHAUSTIVE=1 is set; the run recorded in
profiles/r01_summary.md was: 3 passed in 454 s."""
import os
import threading

import numpy as np
import pytest

import sxtest

pytestmark = pytest.mark.skipif(os.environ.get("SX_EXHAUSTIVE") != "1",
                                reason="exhaustive oracle-vs-reference sweep: set SX_EXHAUSTIVE=1 (about 8 minutes)")

CHUNK = 1 << 26
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


def test_rx_oracle_equals_reference_for_every_word(oracle, ref, nthreads):
    a = np.empty(CHUNK, np.float32)
    b = np.empty(CHUNK, np.float32)
    base = np.arange(CHUNK, dtype=np.uint32)
    for c in range(NCHUNKS):
        words = (base + np.uint32(c * CHUNK)).view(np.int32)

        def work(first, n):
            oracle.sxo_convert_rx_buffer(words.ctypes.data, first // 2, a.ctypes.data, first // 2, n // 2)
            ref.sxref_convert_rx_buffer(words.ctypes.data, first // 2, b.ctypes.data, first // 2, n // 2)

        parallel(work, CHUNK, nthreads)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), c


@pytest.mark.parametrize("slot,other,thr2", [("I", 0.25, sxtest.THR2_DEFAULT), ("Q", -0.75, 0.5625)])
def test_tx_oracle_equals_reference_on_its_whole_defined_domain(oracle, ref, nthreads, slot, other, thr2):
    n = CHUNK // 2                      # frames per chunk
    f = np.empty(2 * n, np.float32)
    a = np.empty(2 * n, np.int32)
    b = np.empty(2 * n, np.int32)
    sweep, fixed = (0, 1) if slot == "I" else (1, 0)
    f[fixed::2] = other
    base = np.arange(n, dtype=np.uint32)
    compared = 0
    for c in range((1 << 32) // n):
        bits = base + np.uint32(c * n)
        f[sweep::2] = bits.view(np.float32)

        def work(first, cnt):
            oracle.sxo_convert_tx_buffer(f.ctypes.data, first, a.ctypes.data, first, cnt, thr2)
            ref.sxref_convert_tx_buffer(f.ctypes.data, first, b.ctypes.data, first, cnt, thr2)

        parallel(work, n, nthreads)
        x = f[sweep::2]
        defined = ~np.isnan(x) & (x < 1.0)
        compared += int(defined.sum())
        pa, pb = a.reshape(-1, 2), b.reshape(-1, 2)
        if not np.array_equal(pa[defined], pb[defined]):
            bad = np.flatnonzero(defined & (pa != pb).any(axis=1))[0]
            pytest.fail(f"TX {slot}: oracle {pa[bad]} != reference {pb[bad]} for input bits {bits[bad]:#x}")
    # every negative float, every non-negative float below 1.0: 2^31 - (NaNs with sign) + 0x3F800000
    assert compared == (1 << 32) - 2 * ((1 << 23) - 1) - ((0x7F800000 - 0x3F800000) + 1)
