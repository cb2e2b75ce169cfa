"""world-size-2 gloo test (CPU) of the multi-GPU host logic: particle sharding, unique-id hand-out, max-over-ranks
timing, and the sum the half-map all-reduce must produce (emulated with a gloo all-reduce on the packed
accumulator layout the library uses: float4 {F.re, F.im, T, 0} per voxel + O + counter)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from thunder_b200 import dist as td
    r, w = td.init("gloo")
    assert (r, w) == (rank, world)
    n = 100003
    a, b = td.shard_range(n, w, r)
    uid = td.share_unique_id(lambda: bytes(range(128)), r, w)
    tmax = td.max_over_ranks([10.0 + r, 5.0 - r])
    # packed accumulators: every rank inserts its own particles; the all-reduce sums them
    rng = np.random.default_rng(7)
    contrib = rng.normal(size=(n, 4)).astype(np.float32)       # one row per particle: what it adds to 1 voxel
    mine = torch.from_numpy(contrib[a:b].sum(0, dtype=np.float64).astype(np.float32))
    dist.all_reduce(mine)
    cnt = td.sum_over_ranks([b - a])
    out.put((rank, a, b, uid, tmax, mine.numpy().tolist(), cnt[0], contrib.sum(0, dtype=np.float64).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, b0, uid0, tmax0, sum0, cnt0, want), (r1, a1, b1, uid1, tmax1, sum1, cnt1, _) = res
    assert (a0, b0, a1, b1) == (0, 50002, 50002, 100003)          # contiguous, first rank takes the extra one
    assert uid0 == uid1 == bytes(range(128))
    assert tmax0 == tmax1 == [11.0, 5.0]                          # max over ranks, element-wise
    assert cnt0 == cnt1 == 100003
    assert np.allclose(sum0, sum1) and np.allclose(sum0, want, rtol=1e-4, atol=1e-2)


def test_shard_range_properties():
    from thunder_b200.dist import shard_range, half_set_of
    for n in (0, 1, 7, 100000, 12345):
        for w in (1, 2, 3, 4, 8):
            rs = [shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    assert [half_set_of(i) for i in range(4)] == [0, 1, 0, 1]
