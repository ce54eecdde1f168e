"""Ceiling check for the host pipeline: chunked H2D + D2H through a 4-slot device ring with no kernels."""
import time, torch
n = 1 << 26  # frames -> 512 MiB per buffer
src = torch.empty(2*n, dtype=torch.int32).pin_memory(); dst = torch.empty(2*n, dtype=torch.int32).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for chunk_lg in (20, 21, 22, 23, 26):
    c = 2 << chunk_lg  # int32 elements per chunk
    K = 4
    dev = [torch.empty(c, dtype=torch.int32, device="cuda") for _ in range(K)]
    nch = (2*n) // c
    def run():
        evs_in = [torch.cuda.Event() for _ in range(nch)]
        evs_out = [torch.cuda.Event() for _ in range(nch)]
        for i in range(nch):
            slot = i % K
            if i >= K: evs_out[i-K].synchronize()
            with torch.cuda.stream(s1):
                dev[slot].copy_(src[i*c:(i+1)*c], non_blocking=True); evs_in[i].record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(evs_in[i])
                dst[i*c:(i+1)*c].copy_(dev[slot], non_blocking=True); evs_out[i].record(s2)
        torch.cuda.synchronize()
    run()
    t0 = time.perf_counter()
    for _ in range(5): run()
    t = (time.perf_counter() - t0) / 5
    print(f"chunk 2^{chunk_lg} frames: {8*n/t/1e9:.1f} GB/s each way ({t*1e3:.2f} ms)")
