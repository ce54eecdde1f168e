"""Ceiling check for the host-buffer leg on N GPUs at once: every rank streams pinned host
memory to its GPU and back (chunked, 4-slot device ring, copy engines only, no kernels) at the
same time.  Launch under torchrun; rank 0 prints one JSON line with GB/s each way per rank.
If N ranks together move no more than bench.py's e2e leg does, the leg sits at what the
platform's host<->device path allows and the limit is outside this library."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 1 << 26  # frames -> 512 MiB per buffer
src = torch.empty(2 * n, dtype=torch.int32).pin_memory()
dst = torch.empty(2 * n, dtype=torch.int32).pin_memory()
src.fill_(rank + 1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
c = 2 << 22  # int32 words per chunk = 2^22 frames, the pipeline's largest chunk
K = 4
dev = [torch.empty(c, dtype=torch.int32, device="cuda") for _ in range(K)]
nch = (2 * n) // c


def both_ways():
    evs_in = [torch.cuda.Event() for _ in range(nch)]
    evs_out = [torch.cuda.Event() for _ in range(nch)]
    for i in range(nch):
        slot = i % K
        if i >= K:
            evs_out[i - K].synchronize()
        with torch.cuda.stream(s1):
            dev[slot].copy_(src[i * c:(i + 1) * c], non_blocking=True)
            evs_in[i].record(s1)
        with torch.cuda.stream(s2):
            s2.wait_event(evs_in[i])
            dst[i * c:(i + 1) * c].copy_(dev[slot], non_blocking=True)
            evs_out[i].record(s2)
    torch.cuda.synchronize()


def one_way(h2d: bool):
    for i in range(nch):
        slot = i % K
        if h2d:
            dev[slot].copy_(src[i * c:(i + 1) * c], non_blocking=True)
        else:
            dst[i * c:(i + 1) * c].copy_(dev[slot], non_blocking=True)
    torch.cuda.synchronize()


def timed(fn, reps=5):
    fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    t = (time.perf_counter() - t0) / reps
    mine = torch.tensor([8 * n / t / 1e9], dtype=torch.float64, device="cuda")
    if world == 1:
        return [float(mine)]
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    return [round(float(x), 1) for x in every]


out = {"n_gpus": world, "bytes_per_buffer": 8 * n, "chunk_frames": c // 2,
       "both_ways_gbs_each_way_per_rank": timed(both_ways),
       "h2d_only_gbs_per_rank": timed(lambda: one_way(True)),
       "d2h_only_gbs_per_rank": timed(lambda: one_way(False))}
out["both_ways_total_each_way"] = round(sum(out["both_ways_gbs_each_way_per_rank"]), 1)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
