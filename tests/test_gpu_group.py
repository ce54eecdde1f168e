"""-m gpu: the group device (SoapySXB200Group: N front-ends, one conversion per period) against
N independent driver=sx devices driven one call at a time -- the unmodified reference driver
when oracle/_ref is present.  BASELINE config 4 with frames that really come from N host-side
ALSA stand-ins: everything an application could observe must agree, member by member."""
import ctypes as C

import numpy as np
import pytest

import sxstream
import sxtest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HAS_TIME = sxstream.HAS_TIME


def comparator():
    if sxstream.REF_LIB.exists():
        return sxstream.Harness(sxstream.REF_LIB), "reference"
    return sxstream.Harness(sxstream.PRODUCT_LIB), "product-device"


def stub_lib():
    """The ALSA stand-in's control surface as linked into the product module."""
    return sxstream.Harness(sxstream.PRODUCT_LIB).lib


@pytest.mark.parametrize("members,period,rate,threshold", [(3, 256, 75000.0, "0"), (70, 256, 300000.0, "0.001"),
                                                           (5, 1000, 32.0e6 / 1536, "0.5")])
def test_group_matches_independent_devices(members, period, rate, threshold):
    from sxxcvr_b200 import plugin
    h, kind = comparator()
    clock_arg = ", clock=32e6" if abs(rate * 1536 - 32.0e6) < 1 else ""
    lat = int(round(768 * 1e9 / rate))
    # (advance_before, mode, offset): read+timed write, read only, untimed write, far / late bursts
    steps = [(0, "rx+", lat)] * 5 + [(70000, "rx+", lat)] + [(0, "rx+", lat)] * 2 + [(0, "rx+", -1_000_000_000)]
    steps += [(0, "rx+", lat), (100000, "untimed", 0), (0, "untimed", 0), (0, "rx+", int(2.5e9)), (0, "none", 0), (3, "rx+", lat)]

    devs = []
    for s in range(members):
        d = h.device("driver=sx" + clock_arg)
        d.set_rate(rate)
        h.lib.sx_alsa_set_capture_seed(d.cap, sxtest.SEED + s)
        rx = d.setup(sxstream.RX, args=f"period={period}")
        tx = d.setup(sxstream.TX, args=f"threshold={threshold}, period={period}")
        assert d.activate(rx) == 0 and d.activate(tx) == 0
        devs.append((d, rx, tx))
    want = []
    for adv, mode, off in steps:
        row = []
        for d, rx, tx in devs:
            if adv:
                d.advance(adv)
            r, fl, t, buf = d.read(rx, period)
            w = None
            if mode == "rx+":
                w = d.write(tx, buf, period, HAS_TIME, t + off)
            elif mode == "untimed":
                w = d.write(tx, buf, period)
            row.append((r, fl, t, buf.copy(), w, d.pointers()))
        want.append(row)

    product = plugin.Harness()
    stub = stub_lib()
    args = f"gpu=0, period={period}, threshold={threshold}" + (clock_arg or "")
    with plugin.Group(product, members, args) as g:
        assert g.period == period
        g.set_rate(rate)
        caps = [g.pcm(i, True) for i in range(members)]
        plays = [g.pcm(i, False) for i in range(members)]
        for i in range(members):
            stub.sx_alsa_set_capture_seed(caps[i], sxtest.SEED + i)
        assert g.activate() == 0
        cf = np.zeros(members * period * 2, np.float32)
        for k, (adv, mode, off) in enumerate(steps):
            if adv:
                for c in caps:
                    stub.sx_alsa_advance(c, adv)
            if mode == "rx+" and k % 2 == 0:          # the fused iteration ...
                rc, rets, wret, t = g.repeat_all(cf.ctypes.data, period, off)
                fl = np.full(members, HAS_TIME, np.int32)
            else:                                     # ... and the two calls it fuses
                rc, rets, fl, t = g.read_all(cf.ctypes.data, period)
                wret = None
                if mode == "rx+":
                    rc2, wret = g.write_all(cf.ctypes.data, period, np.full(members, HAS_TIME), t + off)
                    assert rc2 == 0
                elif mode == "untimed":
                    rc2, wret = g.write_all(cf.ctypes.data, period, np.zeros(members), np.zeros(members))
                    assert rc2 == 0
            assert rc == 0
            got = cf.reshape(members, 2 * period)
            for i in range(members):
                r_, fl_, t_, buf_, w_, ptrs = want[k][i]
                assert (int(rets[i]), int(fl[i]), int(t[i])) == (r_, fl_, t_), (kind, k, i)
                assert np.array_equal(got[i].view(np.uint32), buf_.view(np.uint32)), (kind, k, i)
                if w_ is not None:
                    assert int(wret[i]) == w_, (kind, k, i)
                mine = [stub.sx_alsa_hw_ptr(caps[i]), stub.sx_alsa_appl_ptr(caps[i]), stub.sx_alsa_hw_ptr(plays[i]),
                        stub.sx_alsa_appl_ptr(plays[i])]
                assert mine == list(ptrs), (kind, k, i)
        # what each member's sound card was handed
        for i, (d, rx, tx) in enumerate(devs):
            end = stub.sx_alsa_appl_ptr(plays[i])
            n = min(end, 300000)
            mine = np.empty(2 * n, np.int32)
            stub.sx_alsa_sink_read(plays[i], end - n, n, mine.ctypes.data)
            assert np.array_equal(mine, d.sink(end - n, n)), (kind, i)
            assert mine.any()
        assert g.deactivate() == 0
    for d, _, _ in devs:
        d.close()


@pytest.mark.parametrize("members", [64, 1024])
def test_group_repeater_constant_latency(members):
    """Config 4's invariant on the group: every member's TX block lands exactly 768 frames after
    the RX block it answers, iteration after iteration, while all members share one launch."""
    from sxxcvr_b200 import plugin
    product = plugin.Harness()
    stub = stub_lib()
    rate, period = 75000.0, 256
    lat = int(round(768 * 1e9 / rate))
    with plugin.Group(product, members, "gpu=0, threshold=0") as g:
        g.set_rate(rate)
        assert g.activate() == 0
        for k in range(12):
            rc, rx, tx, t = g.repeat_all(0, period, lat)
            assert rc == 0 and (rx == period).all() and (tx == period).all()
            assert (t == sxstream.Harness(sxstream.PRODUCT_LIB).lib.sxh_ticks_to_time_ns(256 * k, rate)).all()
        for i in (0, members // 2, members - 1):
            assert stub.sx_alsa_appl_ptr(g.pcm(i, True)) == 12 * period
            assert stub.sx_alsa_appl_ptr(g.pcm(i, False)) == 11 * period + 768 + period
        sec = g.bench_repeat(period, lat, 50)
        print(f"group of {members}: {sec / 50 * 1e6:.1f} us per iteration, {sec / 50 / members * 1e6:.3f} us per member")
