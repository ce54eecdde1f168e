"""Is compute-sanitizer initcheck blind to cp.async.bulk stores?  Convert the same block with the
vector schedule and with the bulk (TMA) schedule into fresh cudaMalloc'ed buffers and copy both
back; run under `compute-sanitizer --tool initcheck`."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context

ctx = Context(0)
n = 8192
src = torch.arange(2 * n, dtype=torch.int32).cuda()
for variant in (1, 2, 3):
    ctx.set_option("rx_variant", variant)
    dst = ctx.malloc(16 * n)            # raw cudaMalloc: never written by anything else
    ctx.convert_rx_buffer(src.data_ptr(), 0, dst, 0, n)
    ctx.stream_sync()
    out = np.empty(2 * n, np.float32)
    print(f"variant {variant}: copying the result back", flush=True)
    ctx.memcpy_d2h(out.ctypes.data, dst, 8 * n)
    ctx.stream_sync()
    assert out[5] == np.float32(5 * 2.0**-31)
    ctx.free(dst)
print("done")
