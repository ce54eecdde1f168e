"""Placement of the library's pinned host memory next to the GPU (options numa_local_alloc,
numa_node in include/sxgpu.h).  Placement itself is a performance matter (bench.py reports the
host-buffer leg per rank); here: the options exist, the calling thread's CPU affinity is left
as it was found, and conversions through buffers allocated either way are bit-exact."""
import ctypes
import os

import numpy as np
import pytest

import sxtest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sxxcvr_b200 import Context
    c = Context(0)
    yield c
    c.close()


def test_numa_node_is_reported_and_read_only(ctx):
    node = ctx.get_option("numa_node")
    assert node >= -1
    with pytest.raises(Exception):
        ctx.set_option("numa_node", 0)
    assert ctx.get_option("numa_local_alloc") == 1


@pytest.mark.parametrize("local", [1, 0])
def test_pinned_allocation_keeps_the_callers_affinity_and_converts_exactly(ctx, local):
    oracle = sxtest.load_oracle()
    n = 100003
    before = os.sched_getaffinity(0)
    ctx.set_option("numa_local_alloc", local)
    try:
        src = ctx.malloc_host(8 * n)
        dst = ctx.malloc_host(8 * n)
    finally:
        ctx.set_option("numa_local_alloc", 1)
    assert os.sched_getaffinity(0) == before
    try:
        words = sxtest.rx_uniform(n)
        np.frombuffer((ctypes.c_char * (8 * n)).from_address(src), dtype=np.int32)[:] = words
        ctx.convert_rx_buffer_host(src, 0, dst, 0, n)
        got = np.frombuffer((ctypes.c_char * (8 * n)).from_address(dst), dtype=np.float32).copy()
        assert np.array_equal(got.view(np.uint32), sxtest.oracle_rx(oracle, words).view(np.uint32))
    finally:
        ctx.free_host(src)
        ctx.free_host(dst)
    assert os.sched_getaffinity(0) == before


def test_restricted_caller_affinity_is_never_widened(ctx):
    """A caller confined to one CPU (taskset, a container's cpuset) stays confined."""
    before = os.sched_getaffinity(0)
    one = {sorted(before)[-1]}
    os.sched_setaffinity(0, one)
    try:
        p = ctx.malloc_host(1 << 20)
        assert os.sched_getaffinity(0) == one
        ctx.free_host(p)
    finally:
        os.sched_setaffinity(0, before)
