"""The SoapySDR-compatible Python module (sxxcvr_b200/compat/SoapySDR.py) over the flat
sxh_* view of the driver.

Here (no GPU) it is exercised against the CPU build of the unmodified reference
(oracle/_ref/libsx_ref.so), and -- where /root/reference exists -- by running the
reference's own Python scripts (SoapySX/test/*.py, example/*.py) UNMODIFIED from where they
lie, in a subprocess with the module on PYTHONPATH.  The GPU twin is tests/test_gpu_compat.py.
"""
import json
import os
import re
import signal
import subprocess
import sys
from pathlib import Path

import pytest

from sxstream import REF_LIB, ROOT

COMPAT_DIR = ROOT / "sxxcvr_b200" / "compat"
REFERENCE = Path("/root/reference")

needs_ref_lib = pytest.mark.skipif(not REF_LIB.exists(), reason="oracle/_ref not built")
needs_ref_tree = pytest.mark.skipif(not (REFERENCE / "SoapySX" / "test").is_dir(),
                                    reason="/root/reference is not on this machine")


def run_script(script, *argv, lib=REF_LIB, seconds=120, interrupt_after=None):
    """Run a Python script with `import SoapySDR` resolving to the compat module."""
    env = dict(os.environ)
    env["PYTHONPATH"] = str(COMPAT_DIR) + os.pathsep + env.get("PYTHONPATH", "")
    env["SXSOAPY_LIB"] = str(lib)
    env["MPLBACKEND"] = "Agg"
    proc = subprocess.Popen([sys.executable, str(script), *argv], cwd=str(ROOT), env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        out, err = proc.communicate(timeout=interrupt_after or seconds)
    except subprocess.TimeoutExpired:
        # Scripts that loop for ever are stopped the way a user stops them: Ctrl-C.
        proc.send_signal(signal.SIGINT)
        try:
            out, err = proc.communicate(timeout=30)
        except subprocess.TimeoutExpired:
            proc.kill()
            out, err = proc.communicate()
            if interrupt_after is None:
                raise
    return proc.returncode, out, err


def test_import_fails_loudly_without_a_library(tmp_path):
    env = dict(os.environ, PYTHONPATH=str(COMPAT_DIR), SXSOAPY_LIB=str(tmp_path / "nothing.so"))
    p = subprocess.run([sys.executable, "-c", "import SoapySDR"], env=env, capture_output=True, text=True)
    assert p.returncode != 0
    assert "not found" in p.stderr


def test_module_covers_the_api_the_reference_scripts_use():
    """Every SoapySDR.<name> and device.<method> the reference's scripts touch exists."""
    text = (COMPAT_DIR / "SoapySDR.py").read_text()
    for name in ("SOAPY_SDR_RX", "SOAPY_SDR_TX", "SOAPY_SDR_CF32", "SOAPY_SDR_HAS_TIME", "SOAPY_SDR_INFO",
                 "SOAPY_SDR_DEBUG"):
        assert re.search(r"^%s = " % name, text, re.M), name
    for name in ("setLogLevel", "ticksToTimeNs", "timeNsToTicks"):
        assert re.search(r"^def %s\(" % name, text, re.M), name
    for method in ("setSampleRate", "getSampleRate", "setFrequency", "getFrequency", "setGain", "getGain",
                   "getGainRange", "setupStream", "activateStream", "deactivateStream", "readStream",
                   "writeStream", "getHardwareTime", "readRegisters", "writeRegisters", "closeStream",
                   "getStreamMTU", "listGains", "listAntennas", "setAntenna", "getAntenna"):
        assert re.search(r"^    def %s\(" % method, text, re.M), method
    if (REFERENCE / "example").is_dir():
        used = set()
        for script in list((REFERENCE / "example").glob("*.py")) + list((REFERENCE / "SoapySX" / "test").glob("*.py")):
            src = script.read_text()
            used |= set(re.findall(r"SoapySDR\.([A-Za-z_0-9]+)", src))
            used |= set(re.findall(r"\b(?:dev|device)\.([A-Za-z_0-9]+)\(", src))
        missing = [u for u in sorted(used) if not re.search(r"^\s*(def %s\(|%s = |class %s\b)" % (u, u, u), text, re.M)]
        assert not missing, missing


@needs_ref_lib
def test_timed_repeater_example_runs_against_the_reference_build():
    rc, out, err = run_script(ROOT / "examples" / "timed_repeater.py", "--blocks", "60")
    assert rc == 0, err[-2000:]
    res = json.loads(out.strip().splitlines()[-1])
    assert res["blocks"] == 60 and res["tx_rets"] == [256]
    assert res["first_time_ns"] == 0
    # 256 frames at 75 kHz = 3413333.33 ns: consecutive timestamps step by the floor or the ceiling
    assert set(res["time_steps_ns"]) <= {3413333, 3413334}


@needs_ref_lib
def test_module_in_process_against_the_reference_build(tmp_path):
    """Surface checks that need no script: kwargs forms, enumerate, gains, ranges, registers, errors."""
    code = r'''
import json, numpy as np, SoapySDR
out = {}
out["enumerate"] = SoapySDR.Device.enumerate({"driver": "sx"})
out["enumerate_other"] = SoapySDR.Device.enumerate("driver=hackrf")
d = SoapySDR.Device("driver=sx")
out["keys"] = [d.getDriverKey(), d.getHardwareKey()]
out["formats"] = list(d.getStreamFormats(SoapySDR.SOAPY_SDR_RX, 0))
out["native"] = list(d.getNativeStreamFormat(SoapySDR.SOAPY_SDR_RX, 0))
out["gains_rx"] = list(d.listGains(SoapySDR.SOAPY_SDR_RX, 0))
out["gains_tx"] = list(d.listGains(SoapySDR.SOAPY_SDR_TX, 0))
r = d.getGainRange(SoapySDR.SOAPY_SDR_RX, 0, "LNA")
out["lna_range"] = [r.minimum(), r.maximum(), r.step()]
r = d.getGainRange(SoapySDR.SOAPY_SDR_RX, 0)
out["rx_range"] = [r.minimum(), r.maximum()]
out["gain_at_open"] = [d.getGain(SoapySDR.SOAPY_SDR_RX, 0, "LNA"), d.getGain(SoapySDR.SOAPY_SDR_RX, 0, "PGA"),
                       d.getGain(SoapySDR.SOAPY_SDR_TX, 0, "DAC"), d.getGain(SoapySDR.SOAPY_SDR_TX, 0, "MIXER")]
d.setGain(SoapySDR.SOAPY_SDR_RX, 0, "LNA", 36)
d.setGain(SoapySDR.SOAPY_SDR_RX, 0, "PGA", 10)
out["rx_gain"] = [d.getGain(SoapySDR.SOAPY_SDR_RX, 0, "LNA"), d.getGain(SoapySDR.SOAPY_SDR_RX, 0, "PGA"),
                  d.getGain(SoapySDR.SOAPY_SDR_RX, 0)]
out["antennas"] = list(d.listAntennas(SoapySDR.SOAPY_SDR_RX, 0))
d.setSampleRate(SoapySDR.SOAPY_SDR_RX, 0, 300000.0)
out["rate"] = d.getSampleRate(SoapySDR.SOAPY_SDR_RX, 0)
out["rates"] = len(d.listSampleRates(SoapySDR.SOAPY_SDR_RX, 0))
try:
    d.setupStream(SoapySDR.SOAPY_SDR_RX, SoapySDR.SOAPY_SDR_CS16, [0], {})
    out["cs16"] = "accepted"
except RuntimeError as e:
    out["cs16"] = str(e)
rx = d.setupStream(SoapySDR.SOAPY_SDR_RX, SoapySDR.SOAPY_SDR_CF32)
out["mtu"] = d.getStreamMTU(rx)
d.activateStream(rx)
buf = np.zeros(100, dtype=np.complex64)
try:
    d.readStream(rx, [buf], 101)
    out["short_buffer"] = "accepted"
except ValueError as e:
    out["short_buffer"] = "rejected"
res = d.readStream(rx, [buf], 100)
out["read"] = [res.ret, res.flags, res.timeNs, str(res)]
out["nonzero"] = bool(np.any(buf != 0))
out["ticks"] = [SoapySDR.ticksToTimeNs(75000, 75000.0), SoapySDR.timeNsToTicks(10**9, 75000.0)]
d.deactivateStream(rx); d.closeStream(rx); d.close()
print(json.dumps(out))
'''
    script = tmp_path / "compat_probe.py"
    script.write_text(code)
    rc, out, err = run_script(script)
    assert rc == 0, err[-2000:]
    res = json.loads(out.strip().splitlines()[-1])
    assert len(res["enumerate"]) == 1 and res["enumerate"][0]["driver"] == "sx"
    assert res["enumerate_other"] == []
    assert res["keys"] == ["sx", "sx"]
    assert res["formats"] == ["CF32"]
    assert res["native"][0] == "CF32"
    assert res["gains_rx"] == ["LNA", "PGA"] and res["gains_tx"] == ["DAC", "MIXER"]
    # the reference's register defaults as getGain reads them back before any setGain; the GPU
    # driver starts from the same values (tests/test_gpu_compat.py)
    assert res["gain_at_open"] == [48.0, 30.0, 6.0, 28.0]
    assert res["lna_range"][:2] == [0.0, 48.0]
    assert res["rx_gain"][2] == res["rx_gain"][0] + res["rx_gain"][1]
    assert res["rate"] == 300000.0 and res["rates"] == 6   # the rates of the detected crystal (SoapySX.cpp:193-220)
    assert "accepted" not in res["cs16"]
    assert res["short_buffer"] == "rejected"
    assert res["read"][0] == 100 and res["read"][1] & 4 and res["read"][2] == 0
    assert res["read"][3].startswith("ret=100, flags=")
    assert res["nonzero"]
    assert res["ticks"] == [10**9, 75000]


# --- the reference's own scripts, unmodified ---------------------------------------------

@needs_ref_lib
@needs_ref_tree
def test_reference_test_gains_script():
    rc, out, err = run_script(REFERENCE / "SoapySX" / "test" / "test_gains.py")
    assert rc == 0, err[-2000:]
    assert "RX gain range:" in out and "TX gain range:" in out
    rows = re.findall(r"^\s*(-?[\d.]+) ->\s*(-?[\d.]+) \+\s*(-?[\d.]+) =\s*(-?[\d.]+)$", out, re.M)
    assert len(rows) == 100 + 60
    for want, a, b, total in rows:
        assert abs(float(a) + float(b) - float(total)) < 0.051  # the script prints one decimal
    # the distribution never overshoots the request by more than one element step
    # (LNA 12 dB steps rounded into the table, PGA 2 dB, DAC 3 dB, MIXER 2 dB)
    assert float(rows[99][3]) == 78.0          # RX request 89 dB -> table maximum 48 + 30


@needs_ref_lib
@needs_ref_tree
def test_reference_test_script_registers_and_frequency():
    rc, out, err = run_script(REFERENCE / "SoapySX" / "test" / "test.py")
    assert rc == 0, err[-2000:]
    dumps = [l for l in out.splitlines() if l.startswith("00=")]
    assert len(dumps) == 2 and all(len(re.findall(r"[0-9A-F]{2}=[0-9A-F]{2}", d)) == 0x80 for d in dumps)
    assert dumps[1].startswith("00=0F")        # the script's own writeRegisters('', 0, (0x0F,))
    m = re.search(r"getFrequency: ([\d.]+) ([\d.]+)", out)
    rx, tx = float(m.group(1)), float(m.group(2))
    # the tuning word has 38.4 MHz / 2^20 = 36.6 Hz resolution (reference SoapySX.cpp:1195-1241)
    assert abs(rx - 434.0123456789e6) <= 38.4e6 / 2**20 and abs(tx - 434.123456789e6) <= 38.4e6 / 2**20
    assert "Invalid register address" in err  # "writing too many registers ... should result in an error"


@needs_ref_lib
@needs_ref_tree
def test_reference_linked_streams_script():
    rc, out, err = run_script(REFERENCE / "SoapySX" / "test" / "test_linked_streams.py")
    assert rc == 0, err[-2000:]
    rx = re.findall(r"^RX: ret=(-?\d+), flags=(\d+), timeNs=(-?\d+)$", out, re.M)
    tx = re.findall(r"^TX: ret=(-?\d+), flags=(\d+), timeNs=(-?\d+)$", out, re.M)
    assert len(rx) == 40 and len(tx) == 40
    assert all(int(r[0]) == 256 and int(r[1]) & 4 for r in rx)
    assert all(int(t[0]) == 256 for t in tx)
    times = [int(r[2]) for r in rx]
    assert times[0] == 0                        # linked streams start together at sample 0
    assert all(b - a in (3413333, 3413334) for a, b in zip(times, times[1:]))


@needs_ref_lib
@needs_ref_tree
def test_reference_timestamps_script_until_interrupted():
    rc, out, err = run_script(REFERENCE / "SoapySX" / "test" / "test_timestamps.py", interrupt_after=2)
    rows = re.findall(r"RX T:\s*(-?\d+) TB:\s*(-?\d+) A:\s*(-?\d+) D\s*(-?\d+)", out)
    assert len(rows) > 10, (out[-500:], err[-1500:])
    t = [int(r[0]) for r in rows]
    assert all(b - a in (3413333, 3413334) for a, b in zip(t, t[1:]))
    assert "Uninitializing SoapySX" in err      # Ctrl-C unwinds through the device destructor


@needs_ref_lib
@needs_ref_tree
@pytest.mark.parametrize("script", ["tx_test.py", "linear_repeater.py"])
def test_reference_example_loops_until_interrupted(script):
    pytest.importorskip("scipy")
    rc, out, err = run_script(REFERENCE / "example" / script, interrupt_after=4)
    assert "KeyboardInterrupt" in err, err[-1500:]   # it was still running; nothing else went wrong
    others = [l for l in err.splitlines() if "Error" in l and "KeyboardInterrupt" not in l]
    assert not others, others
    assert "Uninitializing SoapySX" in err


@needs_ref_lib
@needs_ref_tree
def test_reference_plot_rxtx_response_example():
    pytest.importorskip("scipy")
    rc, out, err = run_script(REFERENCE / "example" / "plot_rxtx_response.py", seconds=120)
    assert rc == 0, err[-2000:]
    rows = re.findall(r"^\s*([\d.]+) MHz\s+(-?[\d.]+) dB", out, re.M)
    assert len(rows) >= 20
