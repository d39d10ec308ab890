"""world_size-2 gloo test (CPU) of the batch sharding helpers used for the multi-GPU path."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from shapy_b200 import dist as sdist


def test_shard_bounds():
    assert [sdist.shard_bounds(512, 8, r) for r in range(8)] == [(64 * r, 64 * r + 64) for r in range(8)]
    b = [sdist.shard_bounds(10, 4, r) for r in range(4)]
    assert b == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sdist.shard_bounds(3, 4, 3) == (3, 3)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        per = 3
        full = torch.arange(world * per * 2 * 2 * 2, dtype=torch.float32).view(world * per, 2, 2, 2) if rank == 0 else None
        local = sdist.scatter_images(full, per, (2, 2, 2), torch.device('cpu'))
        expect = torch.arange(world * per * 8, dtype=torch.float32).view(world * per, 2, 2, 2)[rank * per:(rank + 1) * per]
        ok = torch.equal(local, expect)
        # every rank "computes" something position dependent and the results come back in global order
        res = sdist.gather_results({'v': local.sum(dim=(1, 2, 3)).view(per, 1), 'b': local[:, 0, 0]})
        if rank == 0:
            ref = torch.arange(world * per * 8, dtype=torch.float32).view(world * per, 2, 2, 2)
            ok = ok and torch.equal(res['v'], ref.sum(dim=(1, 2, 3)).view(-1, 1)) and torch.equal(res['b'], ref[:, 0, 0])
        else:
            ok = ok and res['v'] is None
        # uint8 scatter used by the configs[4] path (sharded_forward_u8): shards arrive in rank order, byte exact
        nb = 5 * 4 * 4 * 3
        allb = (torch.arange(world * nb) % 251).to(torch.uint8) if rank == 0 else None
        mine = sdist.scatter_bytes(allb, nb, torch.device('cpu'))
        ok = ok and torch.equal(mine, ((torch.arange(world * nb) % 251).to(torch.uint8))[rank * nb:(rank + 1) * nb])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_scatter_gather_gloo_world2():
    world = 2
    ret = mp.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
