"""CPU, world_size 2, gloo: the host-side sharding / gather logic of the multi-GPU render path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG  # noqa: F401  (puts the package on sys.path)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(start: int, count: int) -> torch.Tensor:
    idx = torch.arange(start, start + count)
    return torch.stack([idx % 251, (idx * 7) % 253, (idx // 3) % 255], -1).to(torch.uint8)


def _worker(rank: int, world: int, port: int, totals, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nwx.dist import render_sharded, shard_range
    ok = True
    for total in totals:
        full = render_sharded(total, _fake_render)
        ok &= torch.equal(full, _fake_render(0, total))
        start, count = shard_range(total, rank, world)
        ok &= 0 <= start and start + count <= total
    results[rank] = bool(ok)
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from nwx.dist import shard_range
    for total in (0, 1, 7, 307200, 307201, 921600):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_sharded_render_gathers_full_frame_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, (640 * 480, 1001, 2, 1), results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}


def _grad_worker(rank: int, world: int, port: int, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nwx.training import allreduce_sum_
    flat = torch.full((2, 1000), float(rank + 1))
    scale = allreduce_sum_(flat)
    results[rank] = bool(scale == 0.5 and torch.equal(flat * scale, torch.full((2, 1000), 1.5)))
    dist.destroy_process_group()


def test_gradient_allreduce_mean_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}
    from nwx.training import allreduce_sum_
    t = torch.ones(4)
    assert allreduce_sum_(t) == 1.0 and torch.equal(t, torch.ones(4))      # no process group: identity


def _seed_worker(rank: int, world: int, port: int, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nwx.training import rank_of, rank_seed, world_of
    results[rank] = (rank_seed(7), rank_of(), world_of())
    dist.destroy_process_group()


def test_per_rank_random_streams_gloo_world2():
    """every rank must draw its own batch / jitter / noise -- the RNG key folds the rank in."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_seed_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: (7, 0, 2), 1: (8, 1, 2)}
    from nwx.training import rank_seed
    assert rank_seed(7) == 7                                       # no process group: unshifted


def test_single_process_is_identity():
    from nwx.dist import render_sharded
    assert torch.equal(render_sharded(1000, _fake_render), _fake_render(0, 1000))
