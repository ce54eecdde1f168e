#!/usr/bin/env python
"""L2 eviction policy of the bulk loads/stores x block size, one B200."""
import itertools
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sxxcvr_b200 import Context  # noqa: E402

NAMES = ["first", "normal", "last", "unchanged"]


def main():
    ctx = Context(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    st = side.cuda_stream
    res = []
    for lg in (27, 25):
        n = 1 << lg
        i2s = torch.empty(2 * n, dtype=torch.int32, device="cuda")
        cf = torch.empty(2 * n, dtype=torch.float32, device="cuda")
        out = torch.empty(2 * n, dtype=torch.int32, device="cuda")
        ctx.synth_frames(i2s.data_ptr(), 0, n, 1, st)
        ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)
        for rep in range(2):
            for lp, sp in itertools.product(range(4), range(4)):
                ctx.set_option("bulk_load_policy", lp)
                ctx.set_option("bulk_store_policy", sp)
                row = {"log2_frames": lg, "load": NAMES[lp], "store": NAMES[sp], "rep": rep}
                for name, fn in (("rx", lambda: ctx.convert_rx_buffer(i2s.data_ptr(), 0, cf.data_ptr(), 0, n, st)),
                                 ("tx", lambda: ctx.convert_tx_buffer(cf.data_ptr(), 0, out.data_ptr(), 0, n, 1e-6, st))):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(20):
                        fn()
                    b.record()
                    torch.cuda.synchronize()
                    row[name] = 16 * n / (a.elapsed_time(b) / 20) / 1e6
                res.append(row)
                print(f"2^{lg} load={NAMES[lp]:9s} store={NAMES[sp]:9s} rep{rep}  RX {row['rx']:7.0f}  TX {row['tx']:7.0f} GB/s", flush=True)
        del i2s, cf, out
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/sweep_policy.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
