"""Shared helpers for the test-suite: oracle bindings (the CHECKER, never the product),
seeded input generators, and the KAT tables."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
GOLDEN = ROOT / "tests" / "golden"

SEED = 0x53581255
THR2_DEFAULT = float(np.float32(1.0e-3) * np.float32(1.0e-3))   # bits 0x358637BE (SoapySX.cpp:767-773)
RATES = [32.0e6 / d for d in (1536, 768, 512, 256, 128, 64)] + [38.4e6 / d for d in (1536, 768, 512, 256, 128, 64)]


class OStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("sum", "wsum", "x", "count", "tx_on", "rail")]


def _build_oracle():
    subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)


def load_oracle() -> C.CDLL:
    path = ORACLE_DIR / "libsx_oracle.so"
    src = ORACLE_DIR / "sx_oracle.c"
    if not path.exists() or path.stat().st_mtime < src.stat().st_mtime:
        _build_oracle()
    lib = C.CDLL(str(path))
    P, S = C.c_void_p, C.c_size_t
    lib.sxo_convert_rx_buffer.argtypes = [P, S, P, S, S]
    lib.sxo_convert_tx_buffer.argtypes = [P, S, P, S, S, C.c_float]
    lib.sxo_convert_rx_buffer_cs16.argtypes = [P, S, P, S, S]
    lib.sxo_convert_tx_buffer_cs16.argtypes = [P, S, P, S, S, C.c_float]
    for f in ("sxo_convert_rx_buffer", "sxo_convert_tx_buffer", "sxo_convert_rx_buffer_cs16",
              "sxo_convert_tx_buffer_cs16"):
        getattr(lib, f).restype = None
    lib.sxo_convert_rx_buffer_s16.argtypes = [P, S, P, S, S]
    lib.sxo_convert_rx_buffer_s16.restype = None
    lib.sxo_convert_tx_buffer_s16.argtypes = [P, S, P, S, S, C.c_float]
    lib.sxo_convert_tx_buffer_s16.restype = None
    lib.sxo_ticks_to_time_ns.argtypes = [C.c_longlong, C.c_double]
    lib.sxo_ticks_to_time_ns.restype = C.c_longlong
    lib.sxo_time_ns_to_ticks.argtypes = [C.c_longlong, C.c_double]
    lib.sxo_time_ns_to_ticks.restype = C.c_longlong
    lib.sxo_alsa_sizes.argtypes = [C.c_ulong, C.POINTER(C.c_ulong), C.POINTER(C.c_ulong)]
    lib.sxo_alsa_sizes.restype = None
    lib.sxo_rx_overrun_skip.argtypes = [C.c_long, C.c_ulong, C.c_ulong]
    lib.sxo_rx_overrun_skip.restype = C.c_ulong
    lib.sxo_tx_underrun_forward.argtypes = [C.c_int64, C.c_int64, C.c_ulong]
    lib.sxo_tx_underrun_forward.restype = C.c_int64
    lib.sxo_stats_words.argtypes = [P, S, C.c_uint64, C.POINTER(OStats)]
    lib.sxo_stats_words.restype = None
    lib.sxo_synth_frames.argtypes = [P, C.c_uint64, S, C.c_uint64]
    lib.sxo_synth_frames.restype = None
    return lib


def load_reference():
    path = ORACLE_DIR / "_ref" / "libsx_ref.so"
    if not path.exists():
        if Path("/root/reference/SoapySX/SoapySX.cpp").exists():
            _build_oracle()
        if not path.exists():
            return None
    lib = C.CDLL(str(path))
    P, S = C.c_void_p, C.c_size_t
    lib.sxref_convert_rx_buffer.argtypes = [P, S, P, S, S]
    lib.sxref_convert_rx_buffer.restype = None
    lib.sxref_convert_tx_buffer.argtypes = [P, S, P, S, S, C.c_float]
    lib.sxref_convert_tx_buffer.restype = None
    return lib


# ---- numpy-level wrappers around the oracle -------------------------------------------------
def oracle_rx(lib, words: np.ndarray) -> np.ndarray:
    """int32[2N] I2S words -> float32[2N] (as the reference's convert_rx_buffer)."""
    words = np.ascontiguousarray(words, dtype=np.int32)
    out = np.empty(words.size, np.float32)
    lib.sxo_convert_rx_buffer(words.ctypes.data, 0, out.ctypes.data, 0, words.size // 2)
    return out


def oracle_tx(lib, floats: np.ndarray, thr2: float) -> np.ndarray:
    floats = np.ascontiguousarray(floats, dtype=np.float32)
    out = np.empty(floats.size, np.int32)
    lib.sxo_convert_tx_buffer(floats.ctypes.data, 0, out.ctypes.data, 0, floats.size // 2, thr2)
    return out


def oracle_rx_cs16(lib, words: np.ndarray) -> np.ndarray:
    words = np.ascontiguousarray(words, dtype=np.int32)
    out = np.empty(words.size, np.int16)
    lib.sxo_convert_rx_buffer_cs16(words.ctypes.data, 0, out.ctypes.data, 0, words.size // 2)
    return out


def oracle_tx_cs16(lib, shorts: np.ndarray, thr2: float) -> np.ndarray:
    shorts = np.ascontiguousarray(shorts, dtype=np.int16)
    out = np.empty(shorts.size, np.int32)
    lib.sxo_convert_tx_buffer_cs16(shorts.ctypes.data, 0, out.ctypes.data, 0, shorts.size // 2, thr2)
    return out


def oracle_rx_s16(lib, shorts: np.ndarray) -> np.ndarray:
    shorts = np.ascontiguousarray(shorts, dtype=np.int16)
    out = np.empty(shorts.size, np.float32)
    lib.sxo_convert_rx_buffer_s16(shorts.ctypes.data, 0, out.ctypes.data, 0, shorts.size // 2)
    return out


def oracle_tx_s16(lib, floats: np.ndarray, thr2: float) -> np.ndarray:
    floats = np.ascontiguousarray(floats, dtype=np.float32)
    out = np.empty(floats.size, np.int16)
    lib.sxo_convert_tx_buffer_s16(floats.ctypes.data, 0, out.ctypes.data, 0, floats.size // 2, thr2)
    return out


def ref_rx(lib, words: np.ndarray) -> np.ndarray:
    words = np.ascontiguousarray(words, dtype=np.int32)
    out = np.empty(words.size, np.float32)
    lib.sxref_convert_rx_buffer(words.ctypes.data, 0, out.ctypes.data, 0, words.size // 2)
    return out


def ref_tx(lib, floats: np.ndarray, thr2: float) -> np.ndarray:
    floats = np.ascontiguousarray(floats, dtype=np.float32)
    out = np.empty(floats.size, np.int32)
    lib.sxref_convert_tx_buffer(floats.ctypes.data, 0, out.ctypes.data, 0, floats.size // 2, thr2)
    return out


def oracle_stats(lib, words: np.ndarray, base_index: int = 0):
    w = np.ascontiguousarray(words).view(np.uint32)
    s = OStats()
    lib.sxo_stats_words(w.ctypes.data, w.size, base_index, C.byref(s))
    return (s.sum, s.wsum, s.x, s.count, s.tx_on, s.rail)


def synth_frames(lib, first_frame: int, nframes: int, seed: int = SEED) -> np.ndarray:
    out = np.empty(2 * nframes, np.int32)
    lib.sxo_synth_frames(out.ctypes.data, first_frame, nframes, seed)
    return out


# ---- seeded input generators ---------------------------------------------------------------------
def rx_uniform(nframes: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(-2**31, 2**31, size=2 * nframes, dtype=np.int64).astype(np.int32)


def rx_structured() -> np.ndarray:
    """Boundary words: zero, +-1, the rails, powers of two +-1, the round-to-even edges of
    int->float."""
    vals = [0, 1, -1, 2**31 - 1, -2**31, 0x7FFFFF80, 0x7FFFFFBF, 0x7FFFFFC0, 0x7FFFFFC1, 16777217, 16777216, 16777215,
            -16777217, 0x12345678, 0x87654321 - 2**32, 3, -3, 2, -2]
    for k in range(1, 31):
        vals += [2**k - 1, 2**k, 2**k + 1, -(2**k) - 1, -(2**k), -(2**k) + 1]
    for k in range(24, 31):     # halfway cases for the 24-bit significand
        half = 1 << (k - 24)
        vals += [2**k + half, 2**k + half + 1, 2**k + half - 1, 2**k + 3 * half]
    v = np.array(vals, dtype=np.int64).astype(np.int32)
    if v.size % 2:
        v = np.append(v, np.int32(0))
    return v


ONE_BELOW = np.float32(1.0) - np.float32(2.0**-24)      # 0x3F7FFFFF, largest float < 1


def tx_uniform(nframes: int, seed: int = SEED + 1) -> np.ndarray:
    """Uniform in [-1+2^-24, 1-2^-24]: the domain where the reference is defined C++."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(-1.0, 1.0, size=2 * nframes).astype(np.float32)
    return np.clip(f, -ONE_BELOW, ONE_BELOW)


def tx_gaussian_defined(nframes: int, seed: int = SEED + 2) -> np.ndarray:
    """Gaussian sigma=0.5; values >= 1 are folded just below 1 (defined domain), values <= -1
    are KEPT (the negative clamp is defined: -2^31 fits int32)."""
    rng = np.random.default_rng(seed)
    f = rng.normal(0.0, 0.5, size=2 * nframes).astype(np.float32)
    return np.minimum(f, ONE_BELOW)


def tx_threshold_circle(nframes: int, thr2: float, seed: int = SEED + 3) -> np.ndarray:
    """Points whose |z|^2 lands within a few ulp of thr2: where a fused multiply-add would
    flip the TX-enable bits."""
    rng = np.random.default_rng(seed)
    r = np.sqrt(np.float64(thr2))
    fi = rng.uniform(0.0, r, size=nframes)
    fi32 = fi.astype(np.float32)
    fq = np.sqrt(np.maximum(np.float64(thr2) - fi32.astype(np.float64) ** 2, 0.0)).astype(np.float32)
    ulps = rng.integers(-2, 3, size=nframes).astype(np.int32)
    fq = (fq.view(np.int32) + ulps).view(np.float32)
    sign = rng.integers(0, 2, size=(2, nframes)) * 2 - 1
    out = np.empty(2 * nframes, np.float32)
    out[0::2] = fi32 * sign[0]
    out[1::2] = fq * sign[1]
    return out


def tx_specials() -> np.ndarray:
    """Values outside the reference's defined domain plus awkward in-domain ones.  Checked against
    the C oracle (ARM/saturating semantics) only."""
    f = np.array([1.0, -1.0, 2.0, -2.0, np.inf, -np.inf, np.nan, -np.nan, 1.0000001, -1.0000001,
                  0.0, -0.0, 1e-45, -1e-45, 1.1754944e-38, -1.1754944e-38, 6.98e-10, -6.98e-10,
                  0.99999994, -0.99999994, 0.5, -0.5, 1e-3, 0.000707106781, 3.4e38, -3.4e38,
                  4.656613e-10, -4.656613e-10, 9.313226e-10, -9.313226e-10, 1.8626451e-9, -1.8626451e-9],
                 dtype=np.float32)
    # every ordered pair so each special meets each other as I and as Q
    ii, qq = np.meshgrid(f, f, indexing="ij")
    out = np.empty(2 * ii.size, np.float32)
    out[0::2] = ii.ravel()
    out[1::2] = qq.ravel()
    return out


def in_defined_domain(floats: np.ndarray) -> np.ndarray:
    """Per-frame mask: neither component NaN, both < 1.0."""
    f = floats.reshape(-1, 2)
    ok = ~np.isnan(f) & (f < 1.0)
    return ok.all(axis=1)
