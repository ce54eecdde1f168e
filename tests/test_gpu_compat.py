"""GPU twin of test_compat_soapysdr.py: the same Python programs, through the same `SoapySDR`
module, against the CUDA-backed driver (sxxcvr_b200/lib/libsxsoapy.so) -- and compared with
what the CPU build of the unmodified reference (oracle/_ref) gives for the same program.
Both sit on the same deterministic ALSA stand-in, so timestamps, return values and the CRC
of every received sample must be identical."""
import json

import pytest

from sxstream import PRODUCT_LIB, REF_LIB, ROOT
from test_compat_soapysdr import run_script

pytestmark = pytest.mark.gpu

EXAMPLE = ROOT / "examples" / "timed_repeater.py"


def repeater(lib, *argv):
    rc, out, err = run_script(EXAMPLE, *argv, lib=lib)
    assert rc == 0, err[-3000:]
    return json.loads(out.strip().splitlines()[-1])


@pytest.mark.parametrize("argv", [
    ("--blocks", "80"),
    ("--blocks", "40", "--block", "1000", "--rate", "300000", "--latency", "8192"),
    ("--blocks", "25", "--block", "4096", "--rate", "600000", "--latency", "16384", "--threshold", "0.01"),
    ("--blocks", "40", "--pin"),
], ids=["period_blocks", "ragged_blocks", "large_blocks_threshold", "pinned_caller_buffer"])
def test_timed_repeater_on_the_gpu_matches_the_reference_build(argv):
    got = repeater(PRODUCT_LIB, *argv)
    assert got["blocks"] == int(argv[1]) and got["first_time_ns"] == 0
    assert got["tx_rets"] == [got["block"]]
    if not REF_LIB.exists():
        pytest.skip("oracle/_ref not built: ran on the GPU, nothing to compare with")
    ref_argv = [a for a in argv if a != "--pin"]   # `pin` is a stream argument of the GPU driver only
    want = repeater(REF_LIB, *ref_argv)
    assert got == want


def test_module_surface_on_the_gpu_driver(tmp_path):
    code = r'''
import json, numpy as np, SoapySDR
from SoapySDR import *
out = {}
d = SoapySDR.Device({"driver": "sx"})
out["info"] = d.getHardwareInfo()
out["gain_at_open"] = [d.getGain(SOAPY_SDR_RX, 0, "LNA"), d.getGain(SOAPY_SDR_RX, 0, "PGA"),
                       d.getGain(SOAPY_SDR_TX, 0, "DAC"), d.getGain(SOAPY_SDR_TX, 0, "MIXER")]
out["formats"] = list(d.getStreamFormats(SOAPY_SDR_RX, 0))
d.setSampleRate(SOAPY_SDR_RX, 0, 75000.0); d.setSampleRate(SOAPY_SDR_TX, 0, 75000.0)
rx = d.setupStream(SOAPY_SDR_RX, SOAPY_SDR_CF32, [0], {"link": "1"})
tx = d.setupStream(SOAPY_SDR_TX, SOAPY_SDR_CF32, [0], {"link": "1"})
d.activateStream(rx); d.activateStream(tx)
w = d.writeStream(tx, [np.zeros(1024, dtype=np.complex64)], 1024)
buf = np.zeros(256, dtype=np.complex64)
rows = []
for _ in range(20):
    r = d.readStream(rx, [buf], 256)
    w = d.writeStream(tx, [buf], 256)
    rows.append([r.ret, r.flags, r.timeNs, w.ret])
out["rows"] = rows
out["peak"] = float(np.max(np.abs(buf.view(np.float32))))
d.deactivateStream(rx); d.deactivateStream(tx)
d2 = SoapySDR.Device({"driver": "sx", "cs16": "1"})
out["formats_cs16"] = list(d2.getStreamFormats(SOAPY_SDR_RX, 0))
print(json.dumps(out))
'''
    script = tmp_path / "probe.py"
    script.write_text(code)
    rc, out, err = run_script(script, lib=PRODUCT_LIB)
    assert rc == 0, err[-3000:]
    res = json.loads(out.strip().splitlines()[-1])
    assert "gpu" in res["info"] and int(res["info"]["gpu_sm_count"]) > 0
    assert res["formats"] == ["CF32"] and res["formats_cs16"] == ["CF32", "CS16"]
    # what the reference's register defaults read back before any setGain (init_registers,
    # SoapySX.cpp:146-176: 0x0C = 0x3F -> LNA 48, PGA 30; 0x08 = 0x2E -> DAC 6, MIXER 28)
    assert res["gain_at_open"] == [48.0, 30.0, 6.0, 28.0]
    rows = res["rows"]
    assert all(r[0] == 256 and r[1] & 4 and r[3] == 256 for r in rows)
    assert rows[0][2] == 0
    assert all(b[2] - a[2] in (3413333, 3413334) for a, b in zip(rows, rows[1:]))
    assert 0.0 < res["peak"] <= 1.0
