"""-m gpu: differential fuzzing of the stream path.  Seeded random scripts -- reads and writes of
random sizes, blocking and non-blocking, timed / untimed / late / far-future bursts, clock jumps
that cause overruns and underruns, injected ALSA faults, deactivate/activate cycles, different
periods, rates and link modes -- run on the product device and on the unmodified reference
driver; every observable must agree step by step."""
import errno

import numpy as np
import pytest

import sxstream
import sxtest

pytestmark = pytest.mark.gpu

RX, TX, HAS_TIME = sxstream.RX, sxstream.TX, sxstream.HAS_TIME
FAULT_OPS = {True: (sxstream.OP_AVAIL_DELAY, sxstream.OP_READI, sxstream.OP_FORWARD),
             False: (sxstream.OP_AVAIL_DELAY, sxstream.OP_WRITEI, sxstream.OP_FORWARDABLE, sxstream.OP_FORWARD)}


def make_script(seed):
    rng = np.random.default_rng(seed)
    period = int(rng.choice([0, 64, 256, 1000, 4096]))
    rate_idx = int(rng.integers(0, 12))
    link = bool(rng.random() < 0.15)
    threshold = str(rng.choice(["0", "0.001", "0.3"]))
    ops = []
    for _ in range(int(rng.integers(40, 90))):
        r = rng.random()
        n = int(rng.choice([0, 1, 7, 100, 256, 257, 1000, 4096, 20000]))
        timeout = int(rng.choice([100000, 100000, 100000, 0, -1]))
        if r < 0.35:
            ops.append(("read", n, timeout))
        elif r < 0.70:
            mode = str(rng.choice(["untimed", "rx+", "rx+", "past", "far", "abs0"]))
            ops.append(("write", n, mode, int(rng.integers(0, 40_000_000)), timeout, int(rng.integers(0, 1 << 30))))
        elif r < 0.82:
            ops.append(("advance", int(rng.choice([0, 1, 50, 300, 5000, 66000, 70000, 200000]))))
        elif r < 0.88:
            ops.append(("hwtime",))
        elif r < 0.93:
            capture = bool(rng.random() < 0.5)
            ops.append(("inject", capture, int(rng.choice(FAULT_OPS[capture])), int(rng.choice([-errno.EPIPE, -errno.EIO])),
                        int(rng.integers(0, 3))))
        elif r < 0.97:
            ops.append(("cycle",))
        else:
            ops.append(("free_run", bool(rng.random() < 0.5)))
    return dict(seed=seed, period=period, rate_idx=rate_idx, link=link, threshold=threshold, ops=ops)


def run_script(h, sc):
    tr = []
    clock = "32e6" if sc["rate_idx"] < 6 else "38.4e6"
    with h.device(f"driver=sx, clock={clock}") as d:
        rate = d.rates()[sc["rate_idx"] % 6]
        d.set_rate(rate)
        extra = (f"period={sc['period']}" if sc["period"] else "") + (", link=1" if sc["link"] else "")
        rx = d.setup(RX, args=extra)
        tx = d.setup(TX, args=f"threshold={sc['threshold']}" + (", " + extra if extra else ""))
        tr.append(["activate", d.activate(rx), d.activate(tx), d.mtu(rx)])
        last_t = 0
        active = True
        for op in sc["ops"]:
            try:
                if op[0] == "read":
                    r, fl, t, buf = d.read(rx, op[1], op[2])
                    if r > 0:
                        last_t = t
                    tr.append(["read", r, fl, t if r > 0 else None, sxstream.crc(buf[: 2 * max(r, 0)]), d.pointers()])
                elif op[0] == "write":
                    _, n, mode, off, timeout, seed = op
                    data = sxtest.tx_gaussian_defined(n, seed=seed)
                    if mode == "untimed":
                        w = d.write(tx, data, n, 0, 0, timeout)
                    else:
                        when = {"rx+": last_t + off, "past": last_t - 1_000_000_000 - off, "far": last_t + 2_000_000_000 + off,
                                "abs0": off}[mode]
                        w = d.write(tx, data, n, HAS_TIME, when, timeout)
                    tr.append(["write", mode, w, d.pointers()])
                elif op[0] == "advance":
                    d.advance(op[1])
                elif op[0] == "hwtime":
                    tr.append(["hwtime", d.hw_time()])
                elif op[0] == "inject":
                    d.inject(op[1], op[2], op[3], op[4])
                elif op[0] == "cycle":
                    if active:
                        tr.append(["deactivate", d.deactivate(rx), d.deactivate(tx), d.pointers()])
                    else:
                        tr.append(["activate", d.activate(rx), d.activate(tx)])
                    active = not active
                elif op[0] == "free_run":
                    d.free_run(op[1])
            except sxstream.Threw as e:
                tr.append(["threw", op[0], str(e)])
        end = d.pointers()[3]
        lo = max(0, end - 30000)
        tr.append(["timeline", sxstream.crc(d.sink(lo, end - lo)), sxstream.runs(d.sink_written_mask(lo, min(end - lo, 6000)))])
    return sxstream.normalise(tr)


GOLDEN_FUZZ = sxstream.ROOT / "tests" / "golden" / "stream_fuzz_traces.json"


@pytest.fixture(scope="module")
def harnesses():
    from sxxcvr_b200 import _build
    _build.build_soapy_module()
    ref = sxstream.Harness(sxstream.REF_LIB) if sxstream.REF_LIB.exists() else None
    return sxstream.Harness(sxstream.PRODUCT_LIB), ref


@pytest.fixture(scope="module")
def golden_fuzz():
    import json
    return json.loads(GOLDEN_FUZZ.read_text())


import os

NSEEDS = int(os.environ.get("SX_FUZZ_SEEDS", "40"))      # a soak run raises this (profiles/r01_summary.md)


@pytest.mark.parametrize("seed", range(NSEEDS))
def test_random_script_matches_reference(harnesses, golden_fuzz, seed):
    """Against the live reference driver when oracle/_ref is present, and against the committed
    traces it generated (tests/golden/stream_fuzz_traces.json) for the default seeds."""
    product, ref = harnesses
    sc = make_script(1000 + seed)
    got = run_script(product, sc)
    golden = golden_fuzz.get(str(sc["seed"]))
    if golden is not None:
        assert got == golden, f"seed {sc['seed']}: product differs from the committed reference trace"
    if ref is None:
        if golden is None:
            pytest.skip("no live reference and no committed trace for this seed")
        return
    want = run_script(ref, sc)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, f"seed {sc['seed']} step {i}: product {g} != reference {w}"
    assert len(got) == len(want)
