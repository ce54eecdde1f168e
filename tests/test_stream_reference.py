"""CPU test: the unmodified reference driver, run over the ALSA/SoapySDR stand-ins, still
produces the committed golden traces.  This guards the stand-ins and the scenario scripts; the
product device is held to the same traces in test_gpu_stream.py."""
import json

import pytest

import sxstream

GOLDEN = json.loads(sxstream.GOLDEN_TRACES.read_text())


@pytest.fixture(scope="module")
def ref_harness():
    if not sxstream.REF_LIB.exists():
        pytest.skip("oracle/_ref/libsx_ref.so not built (needs /root/reference)")
    return sxstream.Harness(sxstream.REF_LIB)


@pytest.mark.parametrize("name", sorted(sxstream.SCENARIOS))
def test_reference_reproduces_golden_trace(ref_harness, name):
    assert sxstream.normalise(sxstream.SCENARIOS[name](ref_harness)) == GOLDEN[name]


def test_golden_traces_cover_every_scenario():
    assert set(GOLDEN) == set(sxstream.SCENARIOS)


def test_golden_repeater_has_constant_latency():
    tr = GOLDEN["repeater"]
    blocks = [t for t in tr if t[0] == "blk"]
    assert all(b[1] == 256 and b[2] == 4 and b[5] == 256 for b in blocks)       # ret, HAS_TIME, written
    timeline = [t for t in tr if t[0] == "timeline"][0][1]
    assert timeline["written"] == [[768, 256 * len(blocks)]]                  # TX lands 768 frames after RX


def test_reference_reproduces_golden_fuzz_traces(ref_harness):
    """The 40 default random scripts of tests/test_gpu_stream_fuzz.py, reference side only."""
    import test_gpu_stream_fuzz as fuzz
    golden = json.loads(fuzz.GOLDEN_FUZZ.read_text())
    assert sorted(golden) == [str(s) for s in range(1000, 1040)]
    for seed, want in golden.items():
        assert fuzz.run_script(ref_harness, fuzz.make_script(int(seed))) == want, seed
