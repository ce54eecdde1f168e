"""How independent capture blocks and streams are spread over the GPUs of one box.

Every output word of the sample path depends on one input word (RX) or one I/Q pair (TX)
(reference SoapySX.cpp:108-111, :121-136) and streams have independent frame counters (:378),
so the path shards with no exchange step.  The only cross-rank traffic is the gather of a few
64-bit checksum words after the work is done.  Pure host logic: runs on any torch.distributed
backend (NCCL on the box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

MASK64 = (1 << 64) - 1
STATS_FIELDS = ("sum", "wsum", "xor", "count", "tx_on", "rail")


def stream_owner(stream_id: int, world: int) -> int:
    """Stream s lives on rank s mod G, so its bookkeeping stays on one host thread and one GPU."""
    return stream_id % world


def streams_of_rank(nstreams: int, world: int, rank: int) -> List[int]:
    return [s for s in range(nstreams) if stream_owner(s, world) == rank]


def split_block(nframes: int, world: int, rank: int, align_frames: int = 2) -> Tuple[int, int]:
    """(first_frame, count) of this rank's share of one large block.  Cuts fall on multiples of
    `align_frames` (2 frames = 16 bytes) so every shard keeps vector alignment."""
    units = -(-nframes // align_frames)
    per = -(-units // world)
    first = min(rank * per * align_frames, nframes)
    last = min((rank + 1) * per * align_frames, nframes)
    return first, last - first


def combine_stats(parts: Sequence[Sequence[int]]) -> Tuple[int, ...]:
    """Reduce per-shard statistics (see sxgpu_stats in include/sxgpu.h): every field is a sum
    mod 2^64 except the xor, so the order of the shards does not matter."""
    s = w = x = c = t = r = 0
    for p in parts:
        s = (s + p[0]) & MASK64
        w = (w + p[1]) & MASK64
        x ^= p[2]
        c += p[3]
        t += p[4]
        r += p[5]
    return (s, w, x, c, t, r)


def to_signed(values: Sequence[int]) -> List[int]:
    """uint64 -> int64 bit patterns (torch has no uint64 collectives)."""
    return [v - (1 << 64) if v >= (1 << 63) else v for v in values]


def to_unsigned(values: Sequence[int]) -> List[int]:
    return [int(v) & MASK64 for v in values]


def gather_stats(stats: Sequence[int], device=None) -> List[Tuple[int, ...]]:
    """all_gather one rank's six statistics words; returns the list over ranks (on every rank)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [tuple(stats)]
    t = torch.tensor(to_signed(stats), dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [tuple(to_unsigned(o.tolist())) for o in out]
