"""world_size-2 gloo tests (CPU) of the multi-rank host logic."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vista_b200 import parallel as par


def test_shard_arithmetic():
    assert [len(par.shard_units(25, 8, r)) for r in range(8)] == [4, 3, 3, 3, 3, 3, 3, 3]
    assert par.frame_shards(25, 2) == [(0, 13), (13, 25)]
    cover = [t for a, b in par.frame_shards(25, 8) for t in range(a, b)]
    assert cover == list(range(25))
    assert par.halo_neighbours(0, 4) == (None, 1) and par.halo_neighbours(3, 4) == (2, None)
    assert [list(par.shard_units(5, 2, r)) for r in range(2)] == [[0, 1, 2], [3, 4]]


def _topology_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vista_b200.modules import _RuntimeOwner
        owner = type("Owner", (_RuntimeOwner,), {})()
        owner._rt_init()
        owner.enable_frame_sharding()
        pair = sorted(dist.get_process_group_ranks(owner.pair_group))
        frames = sorted(dist.get_process_group_ranks(owner._shard_group))
        # the pair exchange: lower rank = unconditional half comes first in the gathered buffer
        mine = torch.full((2,), float(rank))
        both = torch.empty(4)
        dist.all_gather_into_tensor(both, mine, group=owner.pair_group)
        q.put((rank, owner.cfg_half, owner._frame_world, pair, frames, both.tolist()))
    finally:
        dist.destroy_process_group()


def test_cfg_split_topology_four_ranks():
    """modules.enable_frame_sharding with an even world: ranks [0, W/2) own the unconditional half, [W/2, W) the
    conditional half; pair (i, i + W/2) shares frames; the pair all-gather orders the halves (u, c)."""
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_topology_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, half, fw, pair, frames, both in res:
        assert half == rank // 2 and fw == 2
        assert pair == [rank % 2, rank % 2 + 2]
        assert frames == [2 * half, 2 * half + 1]
        lo, hi = float(pair[0]), float(pair[1])
        assert both == [lo, lo, hi, hi]


def _halo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vista_b200 import lib
        from vista_b200.sharded import HaloExchange
        nb, T, hw, C = 2, 3, 2, 4
        a = torch.zeros(nb, T, hw, C)
        send_first, send_last = torch.zeros(nb * hw, C), torch.zeros(nb * hw, C)
        recv_prev, recv_next = torch.full((nb * hw, C), -1.0), torch.full((nb * hw, C), -1.0)
        prev, nxt = par.halo_neighbours(rank, world)
        halo = HaloExchange(None, prev, nxt,
                            first=(send_first.view(nb, hw, C), a[:, 0], send_first, recv_prev),
                            last=(send_last.view(nb, hw, C), a[:, T - 1], send_last, recv_next))

        def fill(step):                       # frame t of rank r holds 100 r + 10 t + step
            for t in range(T):
                a[:, t] = 100.0 * rank + 10.0 * t + step
        ok = []

        def check(step):
            if prev is not None:              # the previous rank's LAST frame
                ok.append(bool((recv_prev == 100.0 * prev + 10.0 * (T - 1) + step).all()))
            if nxt is not None:               # the next rank's FIRST frame
                ok.append(bool((recv_next == 100.0 * nxt + step).all()))
        fill(0)
        lib.begin_tape()
        lib.tape_host(halo.start)
        lib.tape_host(halo.wait)
        tape = lib.end_tape()
        check(0)
        for step in (1, 2, 3):                # replays pick up the new data through the same closures
            fill(step)
            lib.replay(tape)
            check(step)
        q.put((rank, ok, len(tape)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_record_and_replay_four_ranks():
    """sharded.HaloExchange under the launch tape, 4 ranks (two interior ranks with both neighbours): the recorded
    start / wait closures move the right boundary frames on every replay."""
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, n in res:
        assert n == 2
        expect = 4 * ((rank > 0) + (rank < world - 1))
        assert len(ok) == expect and all(ok), (rank, ok)
