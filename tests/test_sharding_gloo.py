"""CPU tests of the multi-GPU host logic: partition rules, and a world-size-2 run over gloo in
which each rank checksums its shard and the gathered, combined result equals the checksum of
the whole block."""
import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from sxxcvr_b200 import sharding  # noqa: E402


def test_streams_partition_round_robin():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            mine = sharding.streams_of_rank(37, world, r)
            assert all(sharding.stream_owner(s, world) == r for s in mine)
            seen += mine
        assert sorted(seen) == list(range(37))


@pytest.mark.parametrize("nframes", [0, 1, 2, 3, 255, 256, 65537, 2**27, 2**29 + 5])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_block_split_covers_everything_on_16_byte_cuts(nframes, world):
    pos = 0
    for r in range(world):
        first, count = sharding.split_block(nframes, world, r)
        assert first == pos and first % 2 == 0 or count == 0
        pos += count
    assert pos == nframes


def test_combine_is_order_independent():
    parts = [(2**64 - 1, 5, 0xF0, 10, 1, 0), (2, 2**63, 0x0F, 6, 2, 1), (7, 7, 7, 7, 7, 7)]
    a = sharding.combine_stats(parts)
    b = sharding.combine_stats(parts[::-1])
    assert a == b == ((2**64 - 1 + 2 + 7) % 2**64, (5 + 2**63 + 7) % 2**64, 0xF0 ^ 0x0F ^ 7, 23, 10, 8)
    assert sharding.to_unsigned(sharding.to_signed(a)) == list(a)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nframes, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT / "tests"))
    import sxtest
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = sxtest.load_oracle()
        # Each rank converts its own shard of one block (here with the CPU oracle standing in for
        # the kernel: this test is about the sharding and gather logic, which never sees samples).
        first, count = sharding.split_block(nframes, world, rank)
        frames = sxtest.synth_frames(oracle, first, count)
        out = sxtest.oracle_tx(oracle, sxtest.oracle_rx(oracle, frames), 0.25)
        mine = sxtest.oracle_stats(oracle, out, base_index=2 * first)
        gathered = sharding.gather_stats(mine)
        queue.put((rank, first, count, sharding.combine_stats(gathered)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo_agree_with_the_unsharded_result(oracle):
    import torch.multiprocessing as mp
    import sxtest
    nframes, world = 100003, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = sxtest.oracle_tx(oracle, sxtest.oracle_rx(oracle, sxtest.synth_frames(oracle, 0, nframes)), 0.25)
    want = sxtest.oracle_stats(oracle, whole, 0)
    assert sorted(r[0] for r in results) == [0, 1]
    assert sum(r[2] for r in results) == nframes
    for r in results:
        assert r[3] == want
