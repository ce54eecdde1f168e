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


@pytest.mark.parametrize("ngpus", [1, 2])
def test_one_process_drives_several_gpus(oracle, ngpus):
    """sxgpu_multi_*: one context and one host thread per GPU, a list of blocks split between them
    (SURVEY.md section 8(e)).  With one GPU the same code runs degenerate, so it is exercised on
    every box; with two the halves really run side by side."""
    from sxxcvr_b200 import Multi
    from sxxcvr_b200.capi import Block
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    n, nblocks = (1 << 18) + 5, 7
    with Multi(list(range(ngpus))) as m:
        assert m.lib.sxgpu_multi_size(m.handle) == ngpus
        # host buffers: block b -> GPU b mod G
        words = [sxtest.rx_uniform(n, seed=70 + b) for b in range(nblocks)]
        outs = [np.zeros(2 * n, np.float32) for _ in range(nblocks)]
        m.convert_rx_host([Block(w.ctypes.data, o.ctypes.data, n, 0.0, 0) for w, o in zip(words, outs)])
        for w, o in zip(words, outs):
            assert np.array_equal(o.view(np.uint32), sxtest.oracle_rx(oracle, w).view(np.uint32))
        f = [sxtest.tx_uniform(n, seed=90 + b) for b in range(nblocks)]
        outi = [np.zeros(2 * n, np.int32) for _ in range(nblocks)]
        thr = [sxtest.THR2_DEFAULT if b % 2 else 0.0 for b in range(nblocks)]
        m.convert_tx_host([Block(a.ctypes.data, o.ctypes.data, n, t, 0) for a, o, t in zip(f, outi, thr)])
        for a, o, t in zip(f, outi, thr):
            assert np.array_equal(o, sxtest.oracle_tx(oracle, a, t))
        frames = [m.context(g).counter("frames_rx") for g in range(ngpus)]
        assert sum(frames) == nblocks * n and all(v > 0 for v in frames)
        # device buffers: each block on the GPU its memory is on, one batched launch per GPU
        srcs = [torch.from_numpy(words[b]).to(f"cuda:{b % ngpus}") for b in range(nblocks)]
        dsts = [torch.zeros(2 * n, dtype=torch.float32, device=f"cuda:{b % ngpus}") for b in range(nblocks)]
        torch.cuda.synchronize()
        m.convert_rx_batch([Block(s.data_ptr(), d.data_ptr(), n, 0.0, 0) for s, d in zip(srcs, dsts)])
        m.sync()
        for w, d in zip(words, dsts):
            assert np.array_equal(d.cpu().numpy().view(np.uint32), sxtest.oracle_rx(oracle, w).view(np.uint32))
        # a host pointer in a device list is refused, with a message
        from sxxcvr_b200 import SxGpuError
        with pytest.raises(SxGpuError) as e:
            m.convert_rx_batch([Block(words[0].ctypes.data, dsts[0].data_ptr(), n, 0.0, 0)])
        assert "not device memory" in str(e.value)
