"""Device-level differential testing: drive a SoapySDR "driver=sx" device through the flat sxh_*
harness and record everything an application can observe -- return codes, flags, timestamps,
sample bytes, the playback timeline.  The same scenario functions run against

  * oracle/_ref/libsx_ref.so   the UNMODIFIED reference driver (CPU converters)   -> golden traces
  * sxxcvr_b200/lib/libsxsoapy.so   the product device (CUDA converters)          -> must match

Scenarios follow the call patterns of the reference's own hand-run scripts
(example/linear_repeater.py, example/tx_test.py, SoapySX/test/test_timestamps.py,
SoapySX/test/test_linked_streams.py); the reference has no automated tests to borrow.
"""
from __future__ import annotations

import ctypes as C
import errno
import json
import zlib
from pathlib import Path

import numpy as np

import sxtest

ROOT = Path(__file__).resolve().parent.parent
REF_LIB = ROOT / "oracle" / "_ref" / "libsx_ref.so"
PRODUCT_LIB = ROOT / "sxxcvr_b200" / "lib" / "libsxsoapy.so"
GOLDEN_TRACES = ROOT / "tests" / "golden" / "stream_traces.json"

RX, TX = 1, 0                       # SOAPY_SDR_RX / SOAPY_SDR_TX
HAS_TIME = 1 << 2
THREW = -1000
OP_AVAIL_DELAY, OP_READI, OP_WRITEI, OP_FORWARD, OP_FORWARDABLE, OP_START = range(6)
ONE_BELOW = float(sxtest.ONE_BELOW)


class Threw(Exception):
    pass


class Harness:
    def __init__(self, path):
        self.lib = lib = C.CDLL(str(path))
        P, S, LL = C.c_void_p, C.c_size_t, C.c_longlong
        sig = {
            "sxh_last_error": (C.c_char_p, []),
            "sxh_set_log_level": (None, [C.c_int]),
            "sxh_enumerate": (C.c_char_p, [C.c_char_p]),
            "sxh_make": (C.c_int, [C.c_char_p, C.POINTER(P)]),
            "sxh_unmake": (C.c_int, [P]),
            "sxh_pcm": (P, [P, C.c_int]),
            "sxh_setup_stream": (P, [P, C.c_int, C.c_char_p, C.c_char_p]),
            "sxh_close_stream": (C.c_int, [P, P]),
            "sxh_activate": (C.c_int, [P, P, C.c_int, LL, S]),
            "sxh_deactivate": (C.c_int, [P, P, C.c_int, LL]),
            "sxh_mtu": (C.c_long, [P, P]),
            "sxh_read": (C.c_int, [P, P, P, S, C.POINTER(C.c_int), C.POINTER(LL), C.c_long]),
            "sxh_write": (C.c_int, [P, P, P, S, C.POINTER(C.c_int), LL, C.c_long]),
            "sxh_hardware_time": (C.c_int, [P, C.c_char_p, C.POINTER(LL)]),
            "sxh_has_hardware_time": (C.c_int, [P, C.c_char_p]),
            "sxh_set_sample_rate": (C.c_int, [P, C.c_int, C.c_double]),
            "sxh_get_sample_rate": (C.c_double, [P, C.c_int]),
            "sxh_list_sample_rates": (C.c_int, [P, C.c_int, C.POINTER(C.c_double), C.c_int]),
            "sxh_num_channels": (C.c_int, [P, C.c_int]),
            "sxh_stream_formats": (C.c_char_p, [P, C.c_int]),
            "sxh_native_format": (C.c_char_p, [P, C.c_int, C.POINTER(C.c_double)]),
            "sxh_driver_key": (C.c_char_p, [P]),
            "sxh_hardware_key": (C.c_char_p, [P]),
            "sxh_hardware_info": (C.c_char_p, [P]),
            "sxh_set_frequency": (C.c_int, [P, C.c_int, C.c_double]),
            "sxh_get_frequency": (C.c_double, [P, C.c_int]),
            "sxh_set_gain": (C.c_int, [P, C.c_int, C.c_double]),
            "sxh_write_setting": (C.c_int, [P, C.c_char_p, C.c_char_p]),
            "sxh_ticks_to_time_ns": (LL, [LL, C.c_double]),
            "sxh_time_ns_to_ticks": (LL, [LL, C.c_double]),
            # ALSA stub control surface
            "sx_alsa_advance": (None, [P, C.c_int64]),
            "sx_alsa_set_free_run": (None, [P, C.c_int]),
            "sx_alsa_set_max_transfer": (None, [P, C.c_ulong]),
            "sx_alsa_set_capture_seed": (None, [P, C.c_uint64]),
            "sx_alsa_set_capture_table": (None, [P, P, S]),
            "sx_alsa_set_sink_limit": (None, [P, S]),
            "sx_alsa_sink_read": (S, [P, C.c_int64, S, P]),
            "sx_alsa_sink_clear": (None, [P]),
            "sx_alsa_sink_written": (C.c_int, [P, C.c_int64]),
            "sx_alsa_hw_ptr": (C.c_int64, [P]),
            "sx_alsa_appl_ptr": (C.c_int64, [P]),
            "sx_alsa_buffer_size": (C.c_ulong, [P]),
            "sx_alsa_period_size": (C.c_ulong, [P]),
            "sx_alsa_inject_error": (None, [P, C.c_int, C.c_int, C.c_uint]),
            "sx_alsa_pcm_count": (S, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib.sxh_set_log_level(2)      # CRITICAL: keep the per-call log lines out of test output

    def error(self) -> str:
        return self.lib.sxh_last_error().decode()

    def device(self, args: str = "driver=sx") -> "Dev":
        # The product device takes the crystal as a device argument (clock=...).  The unmodified
        # reference ignores its arguments and probes the chip instead, so for it the fake SX1255
        # in oracle/ref_plugin.cpp is told which crystal to be.
        import os
        import re
        m = re.search(r"clock=([0-9.e+]+)", args)
        if m:
            os.environ["SXREF_CLOCK"] = m.group(1)
        else:
            os.environ.pop("SXREF_CLOCK", None)
        return Dev(self, args)


class Dev:
    def __init__(self, h: Harness, args: str):
        self.h, self.lib = h, h.lib
        p = C.c_void_p()
        if self.lib.sxh_make(args.encode(), C.byref(p)) != 0:
            raise Threw(h.error())
        self.p = p
        self.cap = self.lib.sxh_pcm(p, 1)
        self.play = self.lib.sxh_pcm(p, 0)

    def close(self):
        if self.p:
            self.lib.sxh_unmake(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc == THREW:
            raise Threw(self.h.error())
        return rc

    def setup(self, direction, fmt="CF32", args=""):
        s = self.lib.sxh_setup_stream(self.p, direction, fmt.encode(), args.encode())
        if not s:
            raise Threw(self.h.error())
        return s

    def close_stream(self, s):
        return self._ck(self.lib.sxh_close_stream(self.p, s))

    def activate(self, s):
        return self._ck(self.lib.sxh_activate(self.p, s, 0, 0, 0))

    def deactivate(self, s):
        return self._ck(self.lib.sxh_deactivate(self.p, s, 0, 0))

    def mtu(self, s):
        return self.lib.sxh_mtu(self.p, s)

    def read(self, s, n, timeout_us=100000, buf=None):
        buf = np.full(2 * n, np.float32(-9.0)) if buf is None else buf
        flags, t = C.c_int(-1), C.c_longlong(-1)
        ret = self._ck(self.lib.sxh_read(self.p, s, buf.ctypes.data, n, C.byref(flags), C.byref(t), timeout_us))
        return ret, flags.value, t.value, buf

    def write(self, s, buf, n=None, flags=0, time_ns=0, timeout_us=100000):
        buf = np.ascontiguousarray(buf, dtype=np.float32)
        n = buf.size // 2 if n is None else n
        f = C.c_int(flags)
        return self._ck(self.lib.sxh_write(self.p, s, buf.ctypes.data, n, C.byref(f), time_ns, timeout_us))

    def hw_time(self, what=""):
        t = C.c_longlong(0)
        self._ck(self.lib.sxh_hardware_time(self.p, what.encode(), C.byref(t)))
        return t.value

    def set_rate(self, rate):
        self._ck(self.lib.sxh_set_sample_rate(self.p, RX, rate))
        self._ck(self.lib.sxh_set_sample_rate(self.p, TX, rate))

    def rate(self):
        return self.lib.sxh_get_sample_rate(self.p, RX)

    def rates(self):
        out = (C.c_double * 16)()
        n = self._ck(self.lib.sxh_list_sample_rates(self.p, RX, out, 16))
        return list(out[:n])

    # ---- stub control ---------------------------------------------------------------------------
    def advance(self, frames):
        self.lib.sx_alsa_advance(self.cap, frames)

    def free_run(self, on: bool):
        self.lib.sx_alsa_set_free_run(self.cap, 1 if on else 0)
        self.lib.sx_alsa_set_free_run(self.play, 1 if on else 0)

    def capture_table(self, words: np.ndarray):
        w = np.ascontiguousarray(words, dtype=np.int32)
        self.lib.sx_alsa_set_capture_table(self.cap, w.ctypes.data, w.size // 2)

    def sink(self, position, nframes) -> np.ndarray:
        out = np.empty(2 * nframes, np.int32)
        self.lib.sx_alsa_sink_read(self.play, position, nframes, out.ctypes.data)
        return out

    def sink_written_mask(self, position, nframes) -> np.ndarray:
        return np.array([self.lib.sx_alsa_sink_written(self.play, position + i) for i in range(nframes)], dtype=bool)

    def inject(self, capture: bool, op: int, err: int, skip: int = 0):
        self.lib.sx_alsa_inject_error(self.cap if capture else self.play, op, err, skip)

    def pointers(self):
        L = self.lib
        return [L.sx_alsa_hw_ptr(self.cap), L.sx_alsa_appl_ptr(self.cap), L.sx_alsa_hw_ptr(self.play),
                L.sx_alsa_appl_ptr(self.play)]


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def runs(mask: np.ndarray):
    """[[start, length], ...] of the True runs in a boolean mask."""
    out, start = [], None
    for i, m in enumerate(mask):
        if m and start is None:
            start = i
        if not m and start is not None:
            out.append([start, i - start])
            start = None
    if start is not None:
        out.append([start, len(mask) - start])
    return out


def timeline_digest(dev: Dev, nframes: int):
    return {"written": runs(dev.sink_written_mask(0, nframes)), "crc": crc(dev.sink(0, nframes))}


# =================================================================================================
# Scenarios.  Each takes a Harness and returns a JSON-able trace.
# =================================================================================================
def sc_repeater(h: Harness, rate=75000.0, blocks=24, n=256, latency=768, clock=None):
    """example/linear_repeater.py:50-71 with an identity process()."""
    tr = []
    with h.device("driver=sx" + (f", clock={clock}" if clock else "")) as d:
        d.set_rate(rate)
        rx, tx = d.setup(RX), d.setup(TX, args="threshold=0")
        tr.append(["activate", d.activate(rx), d.activate(tx)])
        lat_ns = int(round(latency * 1e9 / rate))
        for _ in range(blocks):
            r, fl, t, buf = d.read(rx, n)
            w = d.write(tx, buf, n, HAS_TIME, t + lat_ns)
            tr.append(["blk", r, fl, t, crc(buf), w])
        tr.append(["ptrs", d.pointers()])
        tr.append(["timeline", timeline_digest(d, blocks * n + latency + 2 * n)])
        tr.append(["deactivate", d.deactivate(rx), d.deactivate(tx), d.pointers()])
    return tr


def sc_timed_bursts(h: Harness):
    """SoapySX/test/test_timestamps.py:33-49: a 256-frame constant burst 10 ms after every Nth
    block, hardware time read around each read.  Burst level is 1-2^-24, the largest value for
    which the reference is defined C++ (the script itself sends exactly 1.0)."""
    tr = []
    with h.device() as d:
        d.set_rate(75000.0)
        rx, tx = d.setup(RX), d.setup(TX)
        d.activate(rx), d.activate(tx)
        burst = np.zeros(512, np.float32)
        burst[0::2] = ONE_BELOW
        for i in range(40):
            before = d.hw_time()
            r, fl, t, buf = d.read(rx, 256)
            after = d.hw_time()
            w = None
            if i % 8 == 3:
                w = d.write(tx, burst, 256, HAS_TIME, t + 10_000_000)
            tr.append(["blk", before, r, fl, t, after, crc(buf), w])
        tr.append(["timeline", timeline_digest(d, 40 * 256 + 2048)])
        tr.append(["first_burst_words", d.sink(3 * 256 + 750, 4).tolist()])
    return tr


def sc_linked(h: Harness):
    """SoapySX/test/test_linked_streams.py: link=1, prime TX with 1024 frames (that write starts
    both PCMs), then lock-step read/write."""
    tr = []
    with h.device() as d:
        d.set_rate(75000.0)
        rx, tx = d.setup(RX, args="link=1"), d.setup(TX, args="link=1")
        tr.append(["activate", d.activate(rx), d.activate(tx)])
        tr.append(["prime", d.write(tx, np.zeros(2048, np.float32), 1024)])
        for _ in range(40):
            r, fl, t, buf = d.read(rx, 256)
            w = d.write(tx, buf, 256)
            tr.append(["blk", r, fl, t, crc(buf), w])
        tr.append(["ptrs", d.pointers()])
        # stop feeding TX: the linked pair xruns and both directions report it
        d.advance(70000)
        r, fl, t, _ = d.read(rx, 256)
        tr.append(["after_xrun_read", r, fl])
        tr.append(["after_xrun_write", d.write(tx, np.zeros(512, np.float32), 256)])
        tr.append(["deactivate", d.deactivate(rx), d.deactivate(tx)])
    return tr


def sc_overrun(h: Harness):
    """RX overrun, SoapySX.cpp:907-927: pending > ring -> skip whole periods + 2."""
    tr = []
    with h.device() as d:
        d.set_rate(300000.0)
        rx, tx = d.setup(RX, args="period=1024"), d.setup(TX, args="period=1024")
        d.activate(rx), d.activate(tx)
        tr.append(["mtu", d.mtu(rx), d.mtu(tx)])
        for late in (0, 1000, 65536, 65537, 70000, 200000, 64512 + 1024 * 3 + 5):
            d.advance(late)
            r, fl, t, buf = d.read(rx, 1024)
            tr.append(["read", late, r, fl, t, crc(buf), d.pointers()[:2]])
        tr.append(["hw_time", d.hw_time()])
    return tr


def sc_untimed_tx(h: Harness):
    """example/tx_test.py: continuous untimed TX of 4096-frame blocks, then a stall long enough to
    underrun, SoapySX.cpp:1024-1038."""
    tr = []
    with h.device() as d:
        d.set_rate(300000.0)
        tx = d.setup(TX, args="threshold=0")
        rx = d.setup(RX)
        d.activate(tx), d.activate(rx)
        blk = sxtest.tx_uniform(4096, seed=7)
        for i in range(20):
            tr.append(["w", d.write(tx, blk, 4096), d.pointers()[2:]])
        for stall in (100000, 5, 70000):
            d.advance(stall)
            tr.append(["w_after_stall", stall, d.write(tx, blk, 4096), d.pointers()[2:], d.hw_time()])
        tr.append(["timeline_crc", crc(d.sink(0, 20 * 4096))])
    return tr


def sc_late_and_far(h: Harness):
    """Timed TX edge cases: a burst in the past is dropped whole but reported written
    (SoapySX.cpp:1013-1023); a burst further ahead than the ring is reached by forwarding and
    waiting (:1045-1073)."""
    tr = []
    with h.device() as d:
        d.set_rate(75000.0)
        rx, tx = d.setup(RX), d.setup(TX)
        d.activate(rx), d.activate(tx)
        blk = sxtest.tx_uniform(256, seed=11)
        r, fl, t, _ = d.read(rx, 4096)
        tr.append(["read", r, fl, t])
        tr.append(["past", d.write(tx, blk, 256, HAS_TIME, 0), d.pointers()[2:]])
        tr.append(["now", d.write(tx, blk, 256, HAS_TIME, d.hw_time()), d.pointers()[2:]])
        far = d.hw_time() + int(2.5e9)           # 187 500 frames ahead: ~3 rings
        tr.append(["far", d.write(tx, blk, 256, HAS_TIME, far), d.pointers()[2:]])
        tr.append(["far_position", h.lib.sxh_time_ns_to_ticks(far, 75000.0)])
        tr.append(["behind_again", d.write(tx, blk, 256, HAS_TIME, far - 1_000_000_000), d.pointers()[2:]])
        tr.append(["untimed_after", d.write(tx, blk, 256), d.pointers()[2:]])
        tr.append(["hw_time", d.hw_time()])
    return tr


def sc_nonblocking(h: Harness):
    """timeoutUs <= 0 trims to what is there (SoapySX.cpp:934-942, :1076-1085)."""
    tr = []
    with h.device() as d:
        d.set_rate(75000.0)
        rx, tx = d.setup(RX), d.setup(TX, args="threshold=0.5")
        d.activate(rx), d.activate(tx)
        d.free_run(False)
        tr.append(["empty", d.read(rx, 256, 0)[:3]])
        d.advance(100)
        r, fl, t, buf = d.read(rx, 256, 0)
        tr.append(["partial", r, fl, t, crc(buf[:2 * max(r, 0)]), float(buf[2 * max(r, 0)])])
        d.advance(1000)
        r, fl, t, buf = d.read(rx, 256, -5)
        tr.append(["full", r, fl, t, crc(buf)])
        blk = sxtest.tx_uniform(65536 + 100, seed=3)
        tr.append(["tx_trim", d.write(tx, blk, 65536 + 100, 0, 0, 0), d.pointers()[2:]])
        tr.append(["tx_full_ring", d.write(tx, blk, 256, 0, 0, 0), d.pointers()[2:]])
        d.advance(300)
        tr.append(["tx_some_room", d.write(tx, blk, 1000, 0, 0, 0), d.pointers()[2:]])
        tr.append(["sink_head", d.sink(0, 4).tolist()])
    return tr


def sc_errors(h: Harness):
    """Misuse and ALSA failures: which calls throw, which return which code
    (SoapySX.cpp:339-360, :752-764, :815-817, :843-846, :882-894, :981-985)."""
    tr = []

    def threw(fn):
        try:
            fn()
            return None
        except Threw as e:
            return str(e)

    with h.device() as d:
        tr.append(["bad_format", threw(lambda: d.setup(RX, "CS16"))])
        rx = d.setup(RX)
        tr.append(["setup_twice", threw(lambda: d.setup(RX))])
        tx = d.setup(TX, args="threshold=0")
        tr.append(["inactive", d.read(rx, 256)[:2], d.write(tx, np.zeros(512, np.float32), 256)])
        tr.append(["activate", d.activate(rx), d.activate(rx), d.activate(tx)])
        tr.append(["setup_while_running", threw(lambda: d.setup(RX))])
        tr.append(["wrong_direction", threw(lambda: d.read(tx, 256)), threw(lambda: d.write(rx, np.zeros(512, np.float32), 256))])
        tr.append(["bad_time", threw(lambda: d.hw_time("gps"))])
        for op, capture in ((OP_AVAIL_DELAY, True), (OP_READI, True), (OP_FORWARD, True)):
            for err in (-errno.EPIPE, -errno.EIO):
                d.inject(True, op, err)
                if op == OP_FORWARD:
                    d.advance(70000)
                tr.append(["rx_fault", op, err, d.read(rx, 256)[:2]])
        blk = np.zeros(512, np.float32)
        for op in (OP_AVAIL_DELAY, OP_WRITEI, OP_FORWARDABLE, OP_FORWARD):
            for err in (-errno.EPIPE, -errno.EIO):
                d.inject(False, op, err)
                timed = op in (OP_FORWARDABLE, OP_FORWARD)
                tr.append(["tx_fault", op, err,
                           d.write(tx, blk, 256, HAS_TIME if timed else 0, d.hw_time() + 50_000_000 if timed else 0)])
        d.inject(False, OP_AVAIL_DELAY, -errno.EIO)
        tr.append(["hw_time_fault", threw(lambda: d.hw_time())])
        tr.append(["deactivate", d.deactivate(rx), d.deactivate(rx), d.deactivate(tx)])
        tr.append(["after_reset", d.pointers()])
        d.close_stream(rx)
        rx2 = d.setup(RX, args="period=100000")
        tr.append(["period_cap", d.mtu(rx2)])
        tr.append(["bad_threshold", threw(lambda: (d.close_stream(tx), d.setup(TX, args="threshold=abc")))])
    return tr


def sc_identity(h: Harness):
    """Probe surface: SoapySX.cpp:1567-1656."""
    with h.device() as d:
        L = h.lib
        fs = C.c_double(0)
        native = L.sxh_native_format(d.p, RX, C.byref(fs)).decode()
        d.set_rate(75000.0)
        out = {
            "enumerate": L.sxh_enumerate(b"driver=sx").decode(),
            "driver_key": L.sxh_driver_key(d.p).decode(), "hardware_key": L.sxh_hardware_key(d.p).decode(),
            "channels": [L.sxh_num_channels(d.p, RX), L.sxh_num_channels(d.p, TX)],
            "formats": L.sxh_stream_formats(d.p, RX).decode(), "native": [native, fs.value],
            "has_time": [L.sxh_has_hardware_time(d.p, b""), L.sxh_has_hardware_time(d.p, b"x")],
            "rates": d.rates(), "rate": d.rate(),
        }
        bad = []
        for r in (0.0, -1.0, float("nan"), 48000.0, 74000.0, 75001.0):
            try:
                d.set_rate(r)
                bad.append([repr(r), None, d.rate()])
            except Threw as e:
                bad.append([repr(r), str(e), d.rate()])
        out["bad_rates"] = bad
        L.sxh_set_frequency(d.p, RX, 432.55e6)
        L.sxh_set_frequency(d.p, TX, 434.55e6)
        out["freq"] = [L.sxh_get_frequency(d.p, RX), L.sxh_get_frequency(d.p, TX)]
    return [["identity", out]]


def sc_all_rates(h: Harness):
    """Constant RX->TX latency at every legal rate: the TX block written at
    rx.timeNs + round(768e9/rate) lands exactly 768 frames after the RX block."""
    tr = []
    for clock in ("32e6", "38.4e6"):
        with h.device(f"driver=sx, clock={clock}") as d:
            rates = d.rates()
        for rate in rates:
            with h.device(f"driver=sx, clock={clock}") as d:
                d.set_rate(rate)
                rx, tx = d.setup(RX), d.setup(TX, args="threshold=0")
                d.activate(rx), d.activate(tx)
                lat = int(round(768 * 1e9 / rate))
                times, ok = [], True
                for k in range(12):
                    r, fl, t, buf = d.read(rx, 256)
                    d.write(tx, buf, 256, HAS_TIME, t + lat)
                    times.append(t)
                    ok = ok and d.pointers()[3] == 256 * k + 768 + 256
                tr.append(["rate", rate, times[:3], times[-1], ok, runs(d.sink_written_mask(0, 5000))])
    return tr


SCENARIOS = {
    "repeater": sc_repeater,
    "timed_bursts": sc_timed_bursts,
    "linked": sc_linked,
    "overrun": sc_overrun,
    "untimed_tx": sc_untimed_tx,
    "late_and_far": sc_late_and_far,
    "nonblocking": sc_nonblocking,
    "errors": sc_errors,
    "identity": sc_identity,
    "all_rates": sc_all_rates,
}


def normalise(trace):
    """Through JSON and back, so tuples/lists and numpy scalars compare equal to the golden file."""
    return json.loads(json.dumps(trace))
