"""ctypes view of the flat sxh_* harness over a SoapySDR "driver=sx" device (csrc/host/harness_capi.cpp).

The harness talks only to the public SoapySDR::Device interface, so the same binding drives the
product module (lib/libsxsoapy.so: SoapySXB200, CUDA converters) and -- when handed its path --
any other build of the same harness.  bench.py uses it to time readStream/writeStream; nothing
here converts a sample.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import _build

RX, TX = 1, 0            # SOAPY_SDR_RX / SOAPY_SDR_TX
HAS_TIME = 1 << 2
THREW = -1000

_P, _S, _LL = C.c_void_p, C.c_size_t, C.c_longlong

SIGNATURES = {
    "sxh_last_error": (C.c_char_p, []),
    "sxh_set_log_level": (None, [C.c_int]),
    "sxh_make": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "sxh_unmake": (C.c_int, [_P]),
    "sxh_pcm": (_P, [_P, C.c_int]),
    "sxh_setup_stream": (_P, [_P, C.c_int, C.c_char_p, C.c_char_p]),
    "sxh_close_stream": (C.c_int, [_P, _P]),
    "sxh_activate": (C.c_int, [_P, _P, C.c_int, _LL, _S]),
    "sxh_deactivate": (C.c_int, [_P, _P, C.c_int, _LL]),
    "sxh_read": (C.c_int, [_P, _P, _P, _S, C.POINTER(C.c_int), C.POINTER(_LL), C.c_long]),
    "sxh_write": (C.c_int, [_P, _P, _P, _S, C.POINTER(C.c_int), _LL, C.c_long]),
    "sxh_set_sample_rate": (C.c_int, [_P, C.c_int, C.c_double]),
    "sxh_read_setting": (C.c_char_p, [_P, C.c_char_p]),
    "sxh_write_setting": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "sxh_bench_pairs": (C.c_int, [_P, _P, _P, _P, _S, C.c_int, _LL, C.POINTER(C.c_double), _P]),
    "sxh_bench_reads": (C.c_int, [_P, _P, _P, _S, C.c_int, C.POINTER(C.c_double)]),
    "sxh_bench_writes": (C.c_int, [_P, _P, _P, _S, C.c_int, C.POINTER(C.c_double)]),
    "sx_alsa_set_capture_table": (None, [_P, _P, _S]),
    "sx_alsa_set_capture_seed": (None, [_P, C.c_uint64]),
    "sx_alsa_set_sink_limit": (None, [_P, _S]),
    "sx_alsa_sink_read": (_S, [_P, C.c_int64, _S, _P]),
}


class PluginError(RuntimeError):
    pass


class Harness:
    """One loaded build of the harness (default: the product module)."""

    def __init__(self, path: str | Path | None = None):
        path = Path(path) if path else _build.build_soapy_module()
        if not path.exists():
            raise RuntimeError(f"{path} is missing: build it with __graft_entry__.build()")
        self.path = path
        self.lib = lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib.sxh_set_log_level(2)  # keep the driver's per-call log lines out of the way

    def error(self) -> str:
        return self.lib.sxh_last_error().decode()

    def device(self, args: str = "driver=sx") -> "Device":
        return Device(self, args)


class Device:
    def __init__(self, h: Harness, args: str):
        self.h, self.lib = h, h.lib
        p = _P()
        if self.lib.sxh_make(args.encode(), C.byref(p)) != 0:
            raise PluginError(h.error())
        self.p = p
        self.capture = self.lib.sxh_pcm(p, 1)
        self.playback = self.lib.sxh_pcm(p, 0)

    def close(self):
        if self.p:
            self.lib.sxh_unmake(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc, what):
        if rc == THREW:
            raise PluginError(f"{what}: {self.h.error()}")
        return rc

    def set_rate(self, rate: float):
        self._ck(self.lib.sxh_set_sample_rate(self.p, RX, rate), "setSampleRate")
        self._ck(self.lib.sxh_set_sample_rate(self.p, TX, rate), "setSampleRate")

    def setup(self, direction: int, fmt: str = "CF32", args: str = ""):
        s = self.lib.sxh_setup_stream(self.p, direction, fmt.encode(), args.encode())
        if not s:
            raise PluginError(self.h.error())
        return s

    def activate(self, stream):
        return self._ck(self.lib.sxh_activate(self.p, stream, 0, 0, 0), "activateStream")

    def deactivate(self, stream):
        return self._ck(self.lib.sxh_deactivate(self.p, stream, 0, 0), "deactivateStream")

    def read(self, stream, addr: int, n: int, timeout_us: int = 1000000):
        flags, t = C.c_int(0), _LL(0)
        ret = self._ck(self.lib.sxh_read(self.p, stream, addr, n, C.byref(flags), C.byref(t), timeout_us), "readStream")
        return ret, flags.value, t.value

    def write(self, stream, addr: int, n: int, flags: int = 0, time_ns: int = 0, timeout_us: int = 1000000):
        f = C.c_int(flags)
        return self._ck(self.lib.sxh_write(self.p, stream, addr, n, C.byref(f), time_ns, timeout_us), "writeStream")

    def read_setting(self, key: str) -> str:
        return (self.lib.sxh_read_setting(self.p, key.encode()) or b"").decode()

    def write_setting(self, key: str, value: str):
        self._ck(self.lib.sxh_write_setting(self.p, key.encode(), value.encode()), "writeSetting")

    def counter(self, name: str) -> int:
        """A counter of the sxgpu context behind a product device (0 for any other build)."""
        v = self.read_setting("sxgpu." + name)
        return int(v) if v.isdigit() else 0

    # ---- the ALSA stand-in behind the device -----------------------------------------------------
    def capture_table(self, addr: int, nframes: int):
        """Capture frame k = table[k % nframes] (the table is copied)."""
        self.lib.sx_alsa_set_capture_table(self.capture, addr, nframes)

    def sink_limit(self, frames: int):
        self.lib.sx_alsa_set_sink_limit(self.playback, frames)

    def sink(self, position: int, nframes: int, out_addr: int):
        self.lib.sx_alsa_sink_read(self.playback, position, nframes, out_addr)

    # ---- native timing loops (no interpreter between the calls) ----------------------------------
    def bench_pairs(self, rx, tx, addr: int, n: int, iters: int, latency_ns: int, per_iter_addr: int = 0) -> float:
        sec = C.c_double(0)
        rc = self._ck(self.lib.sxh_bench_pairs(self.p, rx, tx, addr, n, iters, latency_ns, C.byref(sec), per_iter_addr),
                      "bench_pairs")
        if rc != 0:
            raise PluginError(f"bench_pairs: a stream call returned {rc}")
        return sec.value

    def bench_reads(self, rx, addr: int, n: int, iters: int) -> float:
        sec = C.c_double(0)
        rc = self._ck(self.lib.sxh_bench_reads(self.p, rx, addr, n, iters, C.byref(sec)), "bench_reads")
        if rc != 0:
            raise PluginError(f"bench_reads: readStream returned {rc}")
        return sec.value

    def bench_writes(self, tx, addr: int, n: int, iters: int) -> float:
        sec = C.c_double(0)
        rc = self._ck(self.lib.sxh_bench_writes(self.p, tx, addr, n, iters, C.byref(sec)), "bench_writes")
        if rc != 0:
            raise PluginError(f"bench_writes: writeStream returned {rc}")
        return sec.value


class Group:
    """N front-ends served as one device: one conversion per period for all of them
    (csrc/host/SoapySXB200Group.hpp, flat sxg_* view).  Product module only."""

    SIGNATURES = {
        "sxg_last_error": (C.c_char_p, []),
        "sxg_create": (C.c_int, [_S, C.c_char_p, C.POINTER(_P)]),
        "sxg_destroy": (C.c_int, [_P]),
        "sxg_size": (_S, [_P]),
        "sxg_period": (_S, [_P]),
        "sxg_set_sample_rate": (C.c_int, [_P, C.c_double]),
        "sxg_activate": (C.c_int, [_P]),
        "sxg_deactivate": (C.c_int, [_P]),
        "sxg_pcm": (_P, [_P, _S, C.c_int]),
        "sxg_read_all": (C.c_int, [_P, _P, _S, _P, _P, _P, C.c_long]),
        "sxg_write_all": (C.c_int, [_P, _P, _S, _P, _P, _P, C.c_long]),
        "sxg_repeat_all": (C.c_int, [_P, _P, _S, _LL, _P, _P, _P, C.c_long]),
        "sxg_bench_repeat": (C.c_int, [_P, _S, _LL, C.c_int, C.POINTER(C.c_double)]),
    }

    def __init__(self, harness: Harness, members: int, args: str = ""):
        import numpy as np
        self.np = np
        self.h, self.lib, self.n = harness, harness.lib, members
        for name, (res, argt) in self.SIGNATURES.items():
            fn = getattr(self.lib, name)
            fn.restype, fn.argtypes = res, argt
        p = _P()
        if self.lib.sxg_create(members, args.encode(), C.byref(p)) != 0:
            raise PluginError(self.lib.sxg_last_error().decode())
        self.p = p
        self.period = self.lib.sxg_period(p)

    def close(self):
        if self.p:
            self.lib.sxg_destroy(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc, what):
        if rc == THREW:
            raise PluginError(f"{what}: {self.lib.sxg_last_error().decode()}")
        return rc

    def set_rate(self, rate):
        self._ck(self.lib.sxg_set_sample_rate(self.p, rate), "setSampleRate")

    def activate(self):
        return self._ck(self.lib.sxg_activate(self.p), "activate")

    def deactivate(self):
        return self._ck(self.lib.sxg_deactivate(self.p), "deactivate")

    def pcm(self, member: int, capture: bool):
        return self.lib.sxg_pcm(self.p, member, 1 if capture else 0)

    def read_all(self, cf32_addr, n, timeout_us=1000000):
        np = self.np
        rets, flags, t = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32), np.zeros(self.n, np.int64)
        rc = self._ck(self.lib.sxg_read_all(self.p, cf32_addr, n, rets.ctypes.data, flags.ctypes.data, t.ctypes.data,
                                            timeout_us), "readAll")
        return rc, rets, flags, t

    def write_all(self, cf32_addr, n, flags, time_ns, timeout_us=1000000):
        np = self.np
        rets = np.zeros(self.n, np.int32)
        f = np.ascontiguousarray(flags, dtype=np.int32)
        t = np.ascontiguousarray(time_ns, dtype=np.int64)
        rc = self._ck(self.lib.sxg_write_all(self.p, cf32_addr, n, f.ctypes.data, t.ctypes.data, rets.ctypes.data,
                                             timeout_us), "writeAll")
        return rc, rets

    def repeat_all(self, cf32_addr, n, offset_ns, timeout_us=1000000):
        np = self.np
        rx, tx, t = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32), np.zeros(self.n, np.int64)
        rc = self._ck(self.lib.sxg_repeat_all(self.p, cf32_addr, n, offset_ns, rx.ctypes.data, tx.ctypes.data,
                                              t.ctypes.data, timeout_us), "repeatAll")
        return rc, rx, tx, t

    def bench_repeat(self, n, offset_ns, iters) -> float:
        sec = C.c_double(0)
        rc = self._ck(self.lib.sxg_bench_repeat(self.p, n, offset_ns, iters, C.byref(sec)), "bench_repeat")
        if rc != 0:
            raise PluginError(f"bench_repeat: a member returned {rc}")
        return sec.value
