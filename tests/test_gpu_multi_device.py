"""-m gpu, needs >= 2 GPUs (skipped otherwise): a context, a bank and a driver=sx device on GPU 1,
alongside GPU 0 in the same process; buffers of the wrong GPU are refused."""
import numpy as np
import pytest

import sxstream
import sxtest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_contexts_on_two_gpus_in_one_process(two_gpus, oracle):
    from sxxcvr_b200 import Context, SxGpuError
    n = (1 << 20) + 3
    with Context(0) as c0, Context(1) as c1:
        assert c0.info().device == 0 and c1.info().device == 1
        outs = []
        for dev, c in ((0, c0), (1, c1)):
            words = sxtest.rx_uniform(n, seed=50 + dev)
            src = torch.from_numpy(words).to(f"cuda:{dev}")
            dst = torch.empty(2 * n, dtype=torch.float32, device=f"cuda:{dev}")
            c.convert_rx_buffer(src.data_ptr(), 0, dst.data_ptr(), 0, n)
            outs.append((c, words, src, dst))
        for c, words, src, dst in outs:
            c.stream_sync()
            assert np.array_equal(dst.cpu().numpy().view(np.uint32), sxtest.oracle_rx(oracle, words).view(np.uint32))
        # a buffer that lives on GPU 0 handed to the context of GPU 1 through the synchronous entry point
        c, words, src0, dst0 = outs[0]
        with pytest.raises(SxGpuError):
            c1.convert_rx_buffer_host(src0.data_ptr(), 0, dst0.data_ptr(), 0, n)


def test_device_on_gpu_1(two_gpus, oracle):
    from sxxcvr_b200 import _build
    _build.build_soapy_module()
    h = sxstream.Harness(sxstream.PRODUCT_LIB)
    with h.device("driver=sx, gpu=1") as d:
        assert "gpu_ordinal=1" in h.lib.sxh_hardware_info(d.p).decode()
        d.set_rate(75000.0)
        rx = d.setup(sxstream.RX)
        d.activate(rx)
        r, fl, t, buf = d.read(rx, 4096)
        assert r == 4096
        assert np.array_equal(buf.view(np.uint32), sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, 0, 4096)).view(np.uint32))
