"""A `SoapySDR` module for Python programs written against SoapySDR's SWIG bindings.

Put this directory on PYTHONPATH and `import SoapySDR` gives the subset of the SWIG module
that the reference's own scripts use (SoapySX/test/*.py, example/*.py in tejeez/sxxcvr):
`Device({'driver': 'sx'})`, sample rate / frequency / gain calls, `setupStream`,
`activateStream`, `readStream` / `writeStream` with numpy buffers returning a result with
`.ret`, `.flags`, `.timeNs`, `getHardwareTime`, `readRegisters` / `writeRegisters`,
`setLogLevel`, `ticksToTimeNs`, and the `SOAPY_SDR_*` constants.

It is a thin ctypes layer over the flat `sxh_*` view of a SoapySDR::Device
(sxxcvr_b200/csrc/host/harness_capi.cpp).  The library is chosen with the environment
variable SXSOAPY_LIB; the default is the CUDA-backed driver module built in this tree
(sxxcvr_b200/lib/libsxsoapy.so).  There is no CPU implementation behind this module: if the
library cannot be loaded the import fails.

On a machine with a real SoapySDR installation this module is not needed: install the
driver module and the stock bindings find it.
"""
import ctypes
import os

import numpy as np

SOAPY_SDR_TX = 0
SOAPY_SDR_RX = 1

SOAPY_SDR_END_BURST = 1 << 1
SOAPY_SDR_HAS_TIME = 1 << 2
SOAPY_SDR_END_ABRUPT = 1 << 3
SOAPY_SDR_ONE_PACKET = 1 << 4
SOAPY_SDR_MORE_FRAGMENTS = 1 << 5
SOAPY_SDR_WAIT_TRIGGER = 1 << 6

SOAPY_SDR_TIMEOUT = -1
SOAPY_SDR_STREAM_ERROR = -2
SOAPY_SDR_CORRUPTION = -3
SOAPY_SDR_OVERFLOW = -4
SOAPY_SDR_NOT_SUPPORTED = -5
SOAPY_SDR_TIME_ERROR = -6
SOAPY_SDR_UNDERFLOW = -7

SOAPY_SDR_FATAL = 1
SOAPY_SDR_CRITICAL = 2
SOAPY_SDR_ERROR = 3
SOAPY_SDR_WARNING = 4
SOAPY_SDR_NOTICE = 5
SOAPY_SDR_INFO = 6
SOAPY_SDR_DEBUG = 7
SOAPY_SDR_TRACE = 8
SOAPY_SDR_SSI = 9

SOAPY_SDR_CF64 = "CF64"
SOAPY_SDR_CF32 = "CF32"
SOAPY_SDR_CS32 = "CS32"
SOAPY_SDR_CS16 = "CS16"
SOAPY_SDR_CS8 = "CS8"
SOAPY_SDR_F32 = "F32"
SOAPY_SDR_S32 = "S32"
SOAPY_SDR_S16 = "S16"

_THREW = -1000

_ERROR_NAMES = {
    SOAPY_SDR_TIMEOUT: "TIMEOUT",
    SOAPY_SDR_STREAM_ERROR: "STREAM_ERROR",
    SOAPY_SDR_CORRUPTION: "CORRUPTION",
    SOAPY_SDR_OVERFLOW: "OVERFLOW",
    SOAPY_SDR_NOT_SUPPORTED: "NOT_SUPPORTED",
    SOAPY_SDR_TIME_ERROR: "TIME_ERROR",
    SOAPY_SDR_UNDERFLOW: "UNDERFLOW",
}

_ELEMENT_BYTES = {"CF64": 16, "CF32": 8, "CS32": 8, "CS16": 4, "CS8": 2, "F32": 4, "S32": 4, "S16": 2}


def _default_library():
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(os.path.dirname(here), "lib", "libsxsoapy.so")


def _load():
    path = os.environ.get("SXSOAPY_LIB") or _default_library()
    if not os.path.exists(path):
        raise ImportError(
            "SoapySDR compat: driver library %s not found (build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or set SXSOAPY_LIB)" % path)
    lib = ctypes.CDLL(path)
    c = ctypes
    P, I, D, S, LL, L, Z, U = (c.c_void_p, c.c_int, c.c_double, c.c_char_p, c.c_longlong, c.c_long,
                               c.c_size_t, c.c_uint)
    sigs = {
        "sxh_last_error": (S, []),
        "sxh_set_log_level": (None, [I]),
        "sxh_enumerate": (S, [S]),
        "sxh_make": (I, [S, c.POINTER(P)]),
        "sxh_unmake": (I, [P]),
        "sxh_setup_stream": (P, [P, I, S, S]),
        "sxh_close_stream": (I, [P, P]),
        "sxh_activate": (I, [P, P, I, LL, Z]),
        "sxh_deactivate": (I, [P, P, I, LL]),
        "sxh_mtu": (L, [P, P]),
        "sxh_read": (I, [P, P, P, Z, c.POINTER(I), c.POINTER(LL), L]),
        "sxh_write": (I, [P, P, P, Z, c.POINTER(I), LL, L]),
        "sxh_hardware_time": (I, [P, S, c.POINTER(LL)]),
        "sxh_has_hardware_time": (I, [P, S]),
        "sxh_set_sample_rate": (I, [P, I, D]),
        "sxh_get_sample_rate": (D, [P, I]),
        "sxh_list_sample_rates": (I, [P, I, c.POINTER(D), I]),
        "sxh_num_channels": (I, [P, I]),
        "sxh_stream_formats": (S, [P, I]),
        "sxh_native_format": (S, [P, I, c.POINTER(D)]),
        "sxh_driver_key": (S, [P]),
        "sxh_hardware_key": (S, [P]),
        "sxh_hardware_info": (S, [P]),
        "sxh_set_frequency": (I, [P, I, D]),
        "sxh_get_frequency": (D, [P, I]),
        "sxh_set_gain_element": (I, [P, I, S, D]),
        "sxh_get_gain": (I, [P, I, S, c.POINTER(D)]),
        "sxh_gain_range": (I, [P, I, S, c.POINTER(D)]),
        "sxh_list_gains": (S, [P, I]),
        "sxh_list_antennas": (S, [P, I]),
        "sxh_set_antenna": (I, [P, I, S]),
        "sxh_get_antenna": (S, [P, I]),
        "sxh_read_registers": (I, [P, S, U, Z, c.POINTER(U)]),
        "sxh_write_registers": (I, [P, S, U, c.POINTER(U), Z]),
        "sxh_read_setting": (S, [P, S]),
        "sxh_write_setting": (I, [P, S, S]),
        "sxh_ticks_to_time_ns": (LL, [LL, D]),
        "sxh_time_ns_to_ticks": (LL, [LL, D]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = a library that is not a driver module
        fn.restype, fn.argtypes = res, args
    return lib


_lib = _load()


def _kwargs_to_text(args):
    if args is None:
        return b""
    if isinstance(args, str):
        return args.encode()
    return ", ".join("%s=%s" % (k, v) for k, v in dict(args).items()).encode()


def _text_to_kwargs(text):
    out = {}
    for item in text.split(","):
        if "=" in item:
            k, v = item.split("=", 1)
            out[k.strip()] = v.strip()
    return out


def _check(rc):
    if rc == _THREW:
        raise RuntimeError(_lib.sxh_last_error().decode())
    return rc


def _split(text):
    text = text.decode() if isinstance(text, bytes) else text
    return tuple(t for t in text.split(",") if t)


def setLogLevel(level):
    _lib.sxh_set_log_level(int(level))


def ticksToTimeNs(ticks, rate):
    return _lib.sxh_ticks_to_time_ns(int(ticks), float(rate))


def timeNsToTicks(timeNs, rate):
    return _lib.sxh_time_ns_to_ticks(int(timeNs), float(rate))


def errToStr(code):
    return _ERROR_NAMES.get(int(code), "UNKNOWN")


def formatToSize(fmt):
    return _ELEMENT_BYTES.get(fmt, 0)


class Range:
    def __init__(self, minimum=0.0, maximum=0.0, step=0.0):
        self._min, self._max, self._step = float(minimum), float(maximum), float(step)

    def minimum(self):
        return self._min

    def maximum(self):
        return self._max

    def step(self):
        return self._step

    def __str__(self):
        text = "%g, %g" % (self._min, self._max)
        if self._step != 0.0:
            text += ", %g" % self._step
        return text

    __repr__ = __str__


class StreamResult:
    """What readStream / writeStream / readStreamStatus return in the SWIG bindings."""

    def __init__(self, ret=0, flags=0, timeNs=0, chanMask=0):
        self.ret, self.flags, self.timeNs, self.chanMask = ret, flags, timeNs, chanMask

    def __str__(self):
        return "ret=%s, flags=%s, timeNs=%s" % (self.ret, self.flags, self.timeNs)

    __repr__ = __str__


class Stream:
    def __init__(self, handle, direction, fmt):
        self.handle, self.direction, self.format = handle, direction, fmt


def _buffer_address(buf, need_bytes, writable):
    """Address of a caller's sample buffer, as SWIG's numpy typemap takes it."""
    if isinstance(buf, np.ndarray):
        if not buf.flags["C_CONTIGUOUS"]:
            raise ValueError("sample buffer must be contiguous")
        if writable and not buf.flags["WRITEABLE"]:
            raise ValueError("readStream needs a writable buffer")
        if buf.nbytes < need_bytes:
            raise ValueError("sample buffer holds %d bytes, the call needs %d" % (buf.nbytes, need_bytes))
        return buf.ctypes.data, buf
    if isinstance(buf, int):  # raw address (pinned or device memory owned by the caller)
        return buf, None
    view = memoryview(buf)
    if writable and view.readonly:
        raise ValueError("readStream needs a writable buffer")
    if view.nbytes < need_bytes:
        raise ValueError("sample buffer holds %d bytes, the call needs %d" % (view.nbytes, need_bytes))
    arr = np.frombuffer(view, dtype=np.uint8)
    return arr.ctypes.data, arr


class Device:
    """SoapySDR.Device: `Device(dict)`, `Device("driver=sx")` or `Device(driver="sx")`."""

    def __init__(self, *args, **kwargs):
        if len(args) > 1:
            raise TypeError("Device takes at most one positional argument")
        text = _kwargs_to_text(args[0] if args else kwargs)
        handle = ctypes.c_void_p()
        _check(_lib.sxh_make(text, ctypes.byref(handle)))
        self._h = handle

    @staticmethod
    def enumerate(args=None):
        text = _lib.sxh_enumerate(_kwargs_to_text(args)).decode()
        return [_text_to_kwargs(t) for t in text.split(";") if t]

    @staticmethod
    def make(*args, **kwargs):
        return Device(*args, **kwargs)

    @staticmethod
    def unmake(device):
        device.close()

    def close(self):
        if getattr(self, "_h", None):
            h, self._h = self._h, None
            _check(_lib.sxh_unmake(h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # identification
    def getDriverKey(self):
        return _lib.sxh_driver_key(self._h).decode()

    def getHardwareKey(self):
        return _lib.sxh_hardware_key(self._h).decode()

    def getHardwareInfo(self):
        return _text_to_kwargs(_lib.sxh_hardware_info(self._h).decode())

    def getNumChannels(self, direction):
        return _check(_lib.sxh_num_channels(self._h, direction))

    # streams
    def getStreamFormats(self, direction, channel=0):
        return _split(_lib.sxh_stream_formats(self._h, direction))

    def getNativeStreamFormat(self, direction, channel=0):
        full = ctypes.c_double(0.0)
        fmt = _lib.sxh_native_format(self._h, direction, ctypes.byref(full)).decode()
        return fmt, full.value

    def setupStream(self, direction, format, channels=(0,), args=None):
        if list(channels) not in ([], [0]):
            raise RuntimeError("setupStream: the device has one channel")
        handle = _lib.sxh_setup_stream(self._h, direction, format.encode(), _kwargs_to_text(args))
        if not handle:
            raise RuntimeError(_lib.sxh_last_error().decode())
        return Stream(handle, direction, format)

    def closeStream(self, stream):
        _check(_lib.sxh_close_stream(self._h, stream.handle))
        stream.handle = None

    def getStreamMTU(self, stream):
        return _lib.sxh_mtu(self._h, stream.handle)

    def activateStream(self, stream, flags=0, timeNs=0, numElems=0):
        return _check(_lib.sxh_activate(self._h, stream.handle, flags, timeNs, numElems))

    def deactivateStream(self, stream, flags=0, timeNs=0):
        return _check(_lib.sxh_deactivate(self._h, stream.handle, flags, timeNs))

    def readStream(self, stream, buffs, numElems, flags=0, timeoutUs=100000):
        if len(buffs) != 1:
            raise ValueError("readStream: one buffer per channel, the device has one channel")
        addr, keep = _buffer_address(buffs[0], numElems * formatToSize(stream.format), True)
        f, t = ctypes.c_int(flags), ctypes.c_longlong(0)
        ret = _check(_lib.sxh_read(self._h, stream.handle, addr, numElems, ctypes.byref(f),
                                   ctypes.byref(t), timeoutUs))
        del keep
        return StreamResult(ret, f.value, t.value)

    def writeStream(self, stream, buffs, numElems, flags=0, timeNs=0, timeoutUs=100000):
        if len(buffs) != 1:
            raise ValueError("writeStream: one buffer per channel, the device has one channel")
        addr, keep = _buffer_address(buffs[0], numElems * formatToSize(stream.format), False)
        f = ctypes.c_int(flags)
        ret = _check(_lib.sxh_write(self._h, stream.handle, addr, numElems, ctypes.byref(f),
                                    timeNs, timeoutUs))
        del keep
        return StreamResult(ret, f.value, 0)

    def readStreamStatus(self, stream, timeoutUs=100000):
        return StreamResult(SOAPY_SDR_NOT_SUPPORTED, 0, 0)

    # time
    def hasHardwareTime(self, what=""):
        return bool(_check(_lib.sxh_has_hardware_time(self._h, what.encode())))

    def getHardwareTime(self, what=""):
        t = ctypes.c_longlong(0)
        _check(_lib.sxh_hardware_time(self._h, what.encode(), ctypes.byref(t)))
        return t.value

    # sample rate
    def setSampleRate(self, direction, channel, rate):
        _check(_lib.sxh_set_sample_rate(self._h, direction, rate))

    def getSampleRate(self, direction, channel):
        return _lib.sxh_get_sample_rate(self._h, direction)

    def listSampleRates(self, direction, channel):
        out = (ctypes.c_double * 64)()
        n = _check(_lib.sxh_list_sample_rates(self._h, direction, out, 64))
        return tuple(out[i] for i in range(min(n, 64)))

    # frequency
    def setFrequency(self, direction, channel, *rest):
        """setFrequency(dir, ch, frequency[, args]) or setFrequency(dir, ch, name, frequency[, args])"""
        values = [r for r in rest if isinstance(r, (int, float))]
        if len(values) != 1:
            raise TypeError("setFrequency(direction, channel, [name,] frequency[, args])")
        _check(_lib.sxh_set_frequency(self._h, direction, float(values[0])))

    def getFrequency(self, direction, channel, name=None):
        return _lib.sxh_get_frequency(self._h, direction)

    # gain
    def listGains(self, direction, channel):
        return _split(_lib.sxh_list_gains(self._h, direction))

    def setGain(self, direction, channel, *rest):
        """setGain(dir, ch, value) or setGain(dir, ch, name, value)"""
        if len(rest) == 1:
            name, value = b"", rest[0]
        elif len(rest) == 2:
            name, value = rest[0].encode(), rest[1]
        else:
            raise TypeError("setGain(direction, channel, [name,] value)")
        _check(_lib.sxh_set_gain_element(self._h, direction, name, float(value)))

    def getGain(self, direction, channel, name=None):
        g = ctypes.c_double(0.0)
        _check(_lib.sxh_get_gain(self._h, direction, (name or "").encode(), ctypes.byref(g)))
        return g.value

    def getGainRange(self, direction, channel, name=None):
        out = (ctypes.c_double * 3)()
        _check(_lib.sxh_gain_range(self._h, direction, (name or "").encode(), out))
        return Range(out[0], out[1], out[2])

    # antenna
    def listAntennas(self, direction, channel):
        return _split(_lib.sxh_list_antennas(self._h, direction))

    def setAntenna(self, direction, channel, name):
        _check(_lib.sxh_set_antenna(self._h, direction, name.encode()))

    def getAntenna(self, direction, channel):
        return _lib.sxh_get_antenna(self._h, direction).decode()

    # registers and settings
    def readRegisters(self, name, addr, length):
        out = (ctypes.c_uint * max(1, length))()
        n = _check(_lib.sxh_read_registers(self._h, name.encode(), addr, length, out))
        return tuple(out[i] for i in range(min(n, length)))

    def writeRegisters(self, name, addr, values):
        values = [int(v) for v in values]
        arr = (ctypes.c_uint * max(1, len(values)))(*values)
        _check(_lib.sxh_write_registers(self._h, name.encode(), addr, arr, len(values)))

    def writeSetting(self, key, value):
        _check(_lib.sxh_write_setting(self._h, key.encode(), str(value).encode()))

    def readSetting(self, key):
        return _lib.sxh_read_setting(self._h, key.encode()).decode()
