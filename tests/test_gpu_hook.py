"""-m gpu: user DSP between RX and TX inside the fused bank iteration (include/sx_hook.cuh,
examples/repeater_hook.cu): the memoryless part of the reference repeater's process() --
s *= 1000; s /= max(|s|, 1); s *= 0.3 (example/linear_repeater.py:87-89, :100-106) -- compiled
into the one-launch iteration, against numpy on the oracle's CF32.

Two comparisons.  (1) Bit-exact against a restatement of the hook's arithmetic in numpy (every
step one correctly rounded IEEE operation; |s| = float32(sqrt(float64(re)^2 + float64(im)^2))).
(2) Against the reference script's literal numpy expression: numpy's np.abs on complex64 is a few
ulp less exact than a correctly rounded hypot (measured: up to 5 ulp), so that comparison carries a
stated tolerance of 8 ulp on the CF32 samples."""
import ctypes as C

import numpy as np
import pytest

import sxtest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HAS_TIME = 1 << 2
PRE, POST = np.float32(1000.0), np.float32(0.3)


@pytest.fixture(scope="module")
def hooklib():
    from sxxcvr_b200 import _build
    lib = C.CDLL(str(_build.build_hook_example()))
    P = C.c_void_p
    lib.sx_example_repeat_clip_gain.argtypes = [P, P, C.c_longlong, C.c_float, C.c_float, P]
    lib.sx_example_repeat_clip_gain.restype = C.c_int
    lib.sx_example_repeat_clip_gain_split.argtypes = [P, P, C.c_uint64, C.c_longlong, C.c_float, C.c_float, P]
    lib.sx_example_repeat_clip_gain_split.restype = C.c_int
    return lib


def clip_gain_exact(cf: np.ndarray) -> np.ndarray:
    """The hook's arithmetic, operation by operation, in numpy."""
    re, im = cf[0::2] * PRE, cf[1::2] * PRE
    mag = np.sqrt(re.astype(np.float64) * re.astype(np.float64) + im.astype(np.float64) * im.astype(np.float64)).astype(np.float32)
    scale = np.float32(1.0) / np.maximum(mag, np.float32(1.0))
    out = np.empty_like(cf)
    out[0::2] = (re * scale) * POST
    out[1::2] = (im * scale) * POST
    return out


def clip_gain_literal(cf: np.ndarray) -> np.ndarray:
    """example/linear_repeater.py:100-106 as written."""
    s = cf.view(np.complex64).copy()
    s *= 1000.0
    s /= np.maximum(np.abs(s), 1.0)
    s *= 0.3
    return s.view(np.float32)


def ulp_distance(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("S,P", [(5, 256), (2100, 256), (17000, 64), (32768 + 3, 64)])
@pytest.mark.parametrize("split", [False, True])
def test_clip_gain_hook_inside_the_fused_iteration(ctx, oracle, hooklib, S, P, split):
    from sxxcvr_b200 import Bank
    rate = 75000.0
    lat = int(round(768 * 1e9 / rate))
    side = torch.cuda.Stream()
    st = side.cuda_stream
    with Bank(ctx, S, P, rate, 1.0e-6, 41) as bank, Bank(ctx, S, P, rate, 1.0e-6, 41) as plain:
        cf = torch.zeros(S * P * 2, dtype=torch.float32, device="cuda")
        cf_plain = torch.zeros_like(cf)
        out = torch.zeros(S * P * 2, dtype=torch.int32, device="cuda")
        for it in range(3):
            with torch.cuda.stream(side):
                if split:
                    rc = hooklib.sx_example_repeat_clip_gain_split(bank.handle, cf.data_ptr(), S * P, lat, PRE, POST, st)
                else:
                    rc = hooklib.sx_example_repeat_clip_gain(bank.handle, cf.data_ptr(), lat, PRE, POST, st)
                assert rc == 0
                bank.drain(0, S, P, out.data_ptr(), st)
                plain.repeat(cf_plain.data_ptr(), lat, st)      # the same iteration with an identity process()
            side.synchronize()
            rx_cf = cf_plain.cpu().numpy()
            for s in (0, S // 2, S - 1):      # the capture itself, against the oracle
                frames = sxtest.synth_frames(oracle, it * P, P, seed=41 + s)
                assert np.array_equal(rx_cf[s * 2 * P:(s + 1) * 2 * P].view(np.uint32),
                                      sxtest.oracle_rx(oracle, frames).view(np.uint32))
            want_cf = clip_gain_exact(rx_cf)
            got_cf = cf.cpu().numpy()
            assert np.array_equal(got_cf.view(np.uint32), want_cf.view(np.uint32)), (it, "hook arithmetic")
            assert ulp_distance(got_cf, clip_gain_literal(rx_cf)).max() <= 8, (it, "literal numpy expression")
            # what was packed for transmission is the TX conversion of the processed block
            assert np.array_equal(out.cpu().numpy(), sxtest.oracle_tx(oracle, want_cf, 1.0e-6)), it
            assert (np.abs(got_cf.view(np.complex64)) <= 0.3 * (1 + 1e-6)).all()       # the clipper clips
            # bookkeeping is that of the plain iteration
            for a, b in zip(bank.positions(st), plain.positions(st)):
                assert np.array_equal(a, b)
            ra, rb = bank.last_read(st), plain.last_read(st)
            assert all(np.array_equal(x, y) for x, y in zip(ra, rb))
