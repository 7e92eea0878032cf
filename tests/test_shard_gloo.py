"""CPU, world_size 2, gloo: the only cross-rank step of the path (final gather of poses / depths)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from super_primitive_b200.shard import gather_ragged, gather_results, owner_of, shard_indices


def test_round_robin_partition():
    for n, w in [(1, 1), (7, 2), (8, 4), (3, 8), (1024, 8)]:
        seen = sorted(i for r in range(w) for i in shard_indices(n, r, w))
        assert seen == list(range(n))
        assert all(owner_of(i, w) == r for r in range(w) for i in shard_indices(n, r, w))


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = shard_indices(n_units, rank, world)
        # unit u has N_u = 3 + u % 4 segments; its "result" is a deterministic function of u
        nloc = max((3 + u % 4 for u in idx), default=1)
        poses = torch.stack([torch.eye(4) * (u + 1) for u in idx]) if idx else torch.zeros((0, 4, 4))
        k = torch.full((len(idx), nloc), float('nan'))
        for i, u in enumerate(idx):
            k[i, :3 + u % 4] = torch.arange(3 + u % 4, dtype=torch.float32) + 10 * u
        cost = torch.tensor([0.5 * u for u in idx], dtype=torch.float32)
        if idx:
            P, K, Cst = gather_results(poses, k, cost, n_units)
        else:       # an empty shard (fewer units than ranks): no batch was ever built on this rank
            P, K, Cst = gather_results(None, None, None, n_units, device="cpu")
        q.put((rank, P.numpy(), K.numpy(), Cst.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [1, 5, 8])
def test_gather_world2(n_units):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + n_units
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, P, K, Cst in outs:
        assert P.shape == (n_units, 4, 4)
        for u in range(n_units):
            assert np.allclose(P[u], np.eye(4) * (u + 1))
            n = 3 + u % 4
            assert np.allclose(K[u, :n], np.arange(n) + 10 * u)
            assert np.all(np.isnan(K[u, n:]))
            assert Cst[u] == pytest.approx(0.5 * u)


def _ragged_worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # unit u (a mapping window) has a result of 5 + 3 * (u % 3) floats
        vecs = [torch.arange(5 + 3 * (u % 3), dtype=torch.float32) + 100 * u for u in shard_indices(n_units, rank, world)]
        out = gather_ragged(vecs, n_units, device="cpu")
        q.put((rank, [o.numpy() for o in out]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [1, 3, 6])
def test_gather_ragged_world2(n_units):
    """Mapping windows have window-dependent result lengths: NaN-padded rows + lengths, one all-gather."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ragged_worker, args=(r, world, 29700 + n_units, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, vecs in outs:
        assert len(vecs) == n_units
        for u, v in enumerate(vecs):
            assert np.array_equal(v, np.arange(5 + 3 * (u % 3), dtype=np.float32) + 100 * u)


def test_gather_ragged_single_process_is_identity():
    vecs = [torch.arange(3.0), torch.arange(5.0)]
    out = gather_ragged(vecs, 2)
    assert all(torch.equal(a, b) for a, b in zip(out, vecs))
