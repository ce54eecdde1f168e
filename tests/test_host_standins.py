"""CPU tests of the hardware/framework stand-ins that both the product device and the oracle
build of the reference compile against: the ALSA PCM model (csrc/shim/alsa_stub.cpp) and the
minimal SoapySDR layer (csrc/shim/soapy_shim.cpp)."""
import ctypes as C
import errno

import numpy as np
import pytest

import sxtest

P, UL, L = C.c_void_p, C.c_ulong, C.c_long
CAPTURE, PLAYBACK = 1, 0
PREPARED, RUNNING, XRUN, SETUP = 2, 3, 4, 1
S32_LE, RW_INTERLEAVED = 10, 3


@pytest.fixture(scope="module")
def lib():
    from sxxcvr_b200 import _build
    lib = C.CDLL(str(_build.build_soapy_module()))
    sig = {
        "snd_pcm_open": (C.c_int, [C.POINTER(P), C.c_char_p, C.c_int, C.c_int]),
        "snd_pcm_close": (C.c_int, [P]), "snd_pcm_prepare": (C.c_int, [P]), "snd_pcm_reset": (C.c_int, [P]),
        "snd_pcm_start": (C.c_int, [P]), "snd_pcm_drop": (C.c_int, [P]), "snd_pcm_state": (C.c_int, [P]),
        "snd_pcm_link": (C.c_int, [P, P]), "snd_pcm_wait": (C.c_int, [P, C.c_int]),
        "snd_pcm_avail_delay": (C.c_int, [P, C.POINTER(L), C.POINTER(L)]),
        "snd_pcm_forwardable": (L, [P]), "snd_pcm_forward": (L, [P, UL]),
        "snd_pcm_readi": (L, [P, P, UL]), "snd_pcm_writei": (L, [P, P, UL]),
        "snd_pcm_hw_params_malloc": (C.c_int, [C.POINTER(P)]), "snd_pcm_hw_params_free": (None, [P]),
        "snd_pcm_hw_params_any": (C.c_int, [P, P]), "snd_pcm_hw_params_set_access": (C.c_int, [P, P, C.c_int]),
        "snd_pcm_hw_params_set_format": (C.c_int, [P, P, C.c_int]),
        "snd_pcm_hw_params_set_channels": (C.c_int, [P, P, C.c_uint]),
        "snd_pcm_hw_params_set_buffer_size_near": (C.c_int, [P, P, C.POINTER(UL)]),
        "snd_pcm_hw_params_set_period_size_near": (C.c_int, [P, P, C.POINTER(UL), P]),
        "snd_pcm_hw_params": (C.c_int, [P, P]),
        "snd_pcm_sw_params_malloc": (C.c_int, [C.POINTER(P)]), "snd_pcm_sw_params_free": (None, [P]),
        "snd_pcm_sw_params_current": (C.c_int, [P, P]), "snd_pcm_sw_params_get_boundary": (C.c_int, [P, C.POINTER(UL)]),
        "snd_pcm_sw_params_set_stop_threshold": (C.c_int, [P, P, UL]), "snd_pcm_sw_params": (C.c_int, [P, P]),
        "sx_alsa_advance": (None, [P, C.c_int64]), "sx_alsa_set_free_run": (None, [P, C.c_int]),
        "sx_alsa_set_max_transfer": (None, [P, UL]), "sx_alsa_set_capture_seed": (None, [P, C.c_uint64]),
        "sx_alsa_set_capture_table": (None, [P, P, C.c_size_t]), "sx_alsa_set_sink_limit": (None, [P, C.c_size_t]),
        "sx_alsa_sink_read": (C.c_size_t, [P, C.c_int64, C.c_size_t, P]), "sx_alsa_sink_written": (C.c_int, [P, C.c_int64]),
        "sx_alsa_sink_clear": (None, [P]), "sx_alsa_hw_ptr": (C.c_int64, [P]), "sx_alsa_appl_ptr": (C.c_int64, [P]),
        "sx_alsa_inject_error": (None, [P, C.c_int, C.c_int, C.c_uint]), "sx_alsa_pcm_count": (C.c_size_t, []),
        "SoapySDR_ticksToTimeNs": (C.c_longlong, [C.c_longlong, C.c_double]),
        "SoapySDR_timeNsToTicks": (C.c_longlong, [C.c_longlong, C.c_double]),
        "SoapySDR_errToStr": (C.c_char_p, [C.c_int]), "SoapySDR_formatToSize": (C.c_size_t, [C.c_char_p]),
        "sxh_enumerate": (C.c_char_p, [C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def open_pcm(lib, direction, period=256, stop_at_buffer=False):
    pcm = P()
    assert lib.snd_pcm_open(C.byref(pcm), b"hw:CARD=SX1255", direction, 0) == 0
    hw = P()
    lib.snd_pcm_hw_params_malloc(C.byref(hw))
    lib.snd_pcm_hw_params_any(pcm, hw)
    assert lib.snd_pcm_hw_params_set_access(pcm, hw, RW_INTERLEAVED) == 0
    assert lib.snd_pcm_hw_params_set_format(pcm, hw, S32_LE) == 0
    assert lib.snd_pcm_hw_params_set_channels(pcm, hw, 2) == 0
    buf, per = UL(65536 // period * period), UL(period)
    lib.snd_pcm_hw_params_set_buffer_size_near(pcm, hw, C.byref(buf))
    lib.snd_pcm_hw_params_set_period_size_near(pcm, hw, C.byref(per), None)
    assert lib.snd_pcm_hw_params(pcm, hw) == 0
    lib.snd_pcm_hw_params_free(hw)
    sw = P()
    lib.snd_pcm_sw_params_malloc(C.byref(sw))
    lib.snd_pcm_sw_params_current(pcm, sw)
    boundary = UL()
    lib.snd_pcm_sw_params_get_boundary(sw, C.byref(boundary))
    lib.snd_pcm_sw_params_set_stop_threshold(pcm, sw, buf.value if stop_at_buffer else boundary.value)
    lib.snd_pcm_sw_params(pcm, sw)
    lib.snd_pcm_sw_params_free(sw)
    return pcm


def avail_delay(lib, pcm):
    a, d = L(), L()
    rc = lib.snd_pcm_avail_delay(pcm, C.byref(a), C.byref(d))
    return rc, a.value, d.value


def test_only_the_sx1255_wire_format_is_accepted(lib):
    pcm = P()
    lib.snd_pcm_open(C.byref(pcm), b"x", CAPTURE, 0)
    hw = P()
    lib.snd_pcm_hw_params_malloc(C.byref(hw))
    lib.snd_pcm_hw_params_any(pcm, hw)
    assert lib.snd_pcm_hw_params_set_format(pcm, hw, 2) == -errno.EINVAL        # S16_LE
    assert lib.snd_pcm_hw_params_set_channels(pcm, hw, 1) == -errno.EINVAL
    assert lib.snd_pcm_hw_params_set_access(pcm, hw, 0) == -errno.EINVAL       # mmap
    big = UL(1 << 20)
    lib.snd_pcm_hw_params_set_buffer_size_near(pcm, hw, C.byref(big))
    assert big.value == 65536                                                   # the Pi's I2S DMA limit
    lib.snd_pcm_hw_params_free(hw)
    lib.snd_pcm_close(pcm)


def test_capture_clock_frames_and_forward(lib, oracle):
    cap = open_pcm(lib, CAPTURE)
    assert lib.snd_pcm_state(cap) == PREPARED
    assert lib.snd_pcm_start(cap) == 0 and lib.snd_pcm_state(cap) == RUNNING
    assert avail_delay(lib, cap) == (0, 0, 0)
    buf = np.zeros(2 * 300, np.int32)
    assert lib.snd_pcm_readi(cap, buf.ctypes.data, 300) == 300                  # blocking read advances the clock
    assert np.array_equal(buf, sxtest.synth_frames(oracle, 0, 300))
    assert (lib.sx_alsa_hw_ptr(cap), lib.sx_alsa_appl_ptr(cap)) == (300, 300)
    lib.sx_alsa_advance(cap, 1000)
    assert avail_delay(lib, cap) == (0, 1000, 1000)
    assert lib.snd_pcm_forwardable(cap) == 1000 and lib.snd_pcm_forward(cap, 5000) == 1000   # clamped to pending
    lib.sx_alsa_set_free_run(cap, 0)
    lib.sx_alsa_advance(cap, 10)
    assert lib.snd_pcm_readi(cap, buf.ctypes.data, 300) == 10                    # short read, no waiting
    assert np.array_equal(buf[:20], sxtest.synth_frames(oracle, 1300, 10))
    lib.sx_alsa_set_free_run(cap, 1)
    lib.sx_alsa_set_max_transfer(cap, 7)
    assert lib.snd_pcm_readi(cap, buf.ctypes.data, 300) == 7
    lib.sx_alsa_advance(cap, 10**6)                                              # stop threshold = boundary: never xruns
    assert lib.snd_pcm_state(cap) == RUNNING and avail_delay(lib, cap)[1] > 65536
    lib.snd_pcm_close(cap)


def test_capture_table_and_seed(lib, oracle):
    cap = open_pcm(lib, CAPTURE)
    table = np.arange(10, dtype=np.int32)
    lib.sx_alsa_set_capture_table(cap, table.ctypes.data, 5)
    lib.snd_pcm_start(cap)
    buf = np.zeros(24, np.int32)
    assert lib.snd_pcm_readi(cap, buf.ctypes.data, 12) == 12
    assert buf.tolist() == (table.tolist() * 3)[:24]
    lib.sx_alsa_set_capture_seed(cap, 42)                                         # back to the generator
    lib.snd_pcm_readi(cap, buf.ctypes.data, 12)
    assert np.array_equal(buf, sxtest.synth_frames(oracle, 12, 12, seed=42))
    lib.snd_pcm_close(cap)


def test_playback_timeline_silence_and_underrun_delay(lib):
    play = open_pcm(lib, PLAYBACK)
    frames = np.arange(2 * 100, dtype=np.int32) + 1
    assert lib.snd_pcm_writei(play, frames.ctypes.data, 100) == 100              # start threshold 1: now running
    assert lib.snd_pcm_state(play) == RUNNING
    assert avail_delay(lib, play) == (0, 65536 - 100, 100)
    assert lib.snd_pcm_forward(play, 50) == 50                                    # a gap: plays as silence
    lib.snd_pcm_writei(play, frames.ctypes.data, 100)
    out = np.zeros(2 * 260, np.int32)
    lib.sx_alsa_sink_read(play, 0, 260, out.ctypes.data)
    assert np.array_equal(out[:200], frames) and not out[200:300].any() and np.array_equal(out[300:500], frames)
    assert not out[500:].any()
    assert [lib.sx_alsa_sink_written(play, p) for p in (0, 99, 100, 149, 150, 249, 250)] == [1, 1, 0, 0, 1, 1, 0]
    lib.sx_alsa_advance(play, 1000)                                               # the DAC overtakes the writer
    rc, avail, delay = avail_delay(lib, play)
    assert (rc, delay) == (0, 250 - 1000) and avail == 65536 + 750                # negative delay = underrun
    lib.sx_alsa_sink_clear(play)
    assert not lib.sx_alsa_sink_written(play, 0)
    lib.snd_pcm_close(play)


def test_blocking_write_waits_for_ring_space(lib):
    play = open_pcm(lib, PLAYBACK)
    lib.sx_alsa_set_sink_limit(play, 1 << 18)
    block = np.ones(2 * 65536, np.int32)
    assert lib.snd_pcm_writei(play, block.ctypes.data, 65536) == 65536           # fills the ring
    assert lib.sx_alsa_hw_ptr(play) == 0
    assert lib.snd_pcm_writei(play, block.ctypes.data, 1000) == 1000             # has to wait 1000 frame times
    assert lib.sx_alsa_hw_ptr(play) == 1000
    lib.sx_alsa_set_free_run(play, 0)
    assert lib.snd_pcm_writei(play, block.ctypes.data, 10) == 0                  # ring full, not allowed to wait
    assert lib.snd_pcm_wait(play, 1) == 0                                         # timed out
    lib.sx_alsa_set_free_run(play, 1)
    assert lib.snd_pcm_wait(play, 1) == 1 and lib.sx_alsa_hw_ptr(play) == 1256   # one period of space
    lib.snd_pcm_close(play)


def test_linked_pair_starts_stops_and_xruns_together(lib):
    cap, play = open_pcm(lib, CAPTURE, stop_at_buffer=True), open_pcm(lib, PLAYBACK, stop_at_buffer=True)
    assert lib.snd_pcm_link(cap, play) == 0 and lib.snd_pcm_link(cap, play) == -errno.EALREADY
    z = np.zeros(2 * 1024, np.int32)
    assert lib.snd_pcm_writei(play, z.ctypes.data, 1024) == 1024                 # first write starts both
    assert lib.snd_pcm_state(cap) == RUNNING and lib.snd_pcm_state(play) == RUNNING
    lib.sx_alsa_advance(play, 500)
    assert lib.sx_alsa_hw_ptr(cap) == 500 and lib.sx_alsa_hw_ptr(play) == 500    # one clock
    lib.sx_alsa_advance(cap, 600)                                                 # playback runs dry at 1024
    assert lib.snd_pcm_state(cap) == XRUN and lib.snd_pcm_state(play) == XRUN
    assert avail_delay(lib, cap)[0] == -errno.EPIPE and lib.snd_pcm_readi(cap, z.ctypes.data, 10) == -errno.EPIPE
    assert lib.snd_pcm_writei(play, z.ctypes.data, 10) == -errno.EPIPE
    assert lib.snd_pcm_drop(cap) == 0 and lib.snd_pcm_state(play) == SETUP       # drop acts on the group
    assert lib.snd_pcm_prepare(play) == 0 and lib.snd_pcm_state(cap) == PREPARED
    assert (lib.sx_alsa_hw_ptr(cap), lib.sx_alsa_appl_ptr(play)) == (0, 0)
    lib.snd_pcm_close(cap)
    lib.snd_pcm_close(play)


def test_fault_injection_is_one_shot_and_skippable(lib):
    cap = open_pcm(lib, CAPTURE)
    lib.snd_pcm_start(cap)
    lib.sx_alsa_inject_error(cap, 0, -errno.EIO, 1)                               # second avail_delay fails
    assert avail_delay(lib, cap)[0] == 0
    assert avail_delay(lib, cap)[0] == -errno.EIO
    assert avail_delay(lib, cap)[0] == 0
    lib.snd_pcm_close(cap)


def test_registry_probe_and_small_helpers(lib):
    assert lib.sxh_enumerate(b"driver=sx") == b"driver=sx, label=sx"
    assert lib.sxh_enumerate(b"driver=other") == b""
    assert lib.SoapySDR_errToStr(-4) == b"OVERFLOW" and lib.SoapySDR_errToStr(-7) == b"UNDERFLOW"
    assert lib.SoapySDR_formatToSize(b"CF32") == 8 and lib.SoapySDR_formatToSize(b"CS16") == 4


@pytest.mark.parametrize("rate", sxtest.RATES)
def test_shim_time_conversion_equals_the_oracle(lib, oracle, rate):
    rng = np.random.default_rng(int(rate))
    for p in [0, 1, 255, 256, 768, 75000, 2**40] + [int(x) for x in rng.integers(0, 2**46, size=300)]:
        assert lib.SoapySDR_ticksToTimeNs(p, rate) == oracle.sxo_ticks_to_time_ns(p, rate)
        assert lib.SoapySDR_timeNsToTicks(p, rate) == oracle.sxo_time_ns_to_ticks(p, rate)
