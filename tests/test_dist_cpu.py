"""Host-side multi-process logic (env sharding, gather to the learner rank, action scatter) on
CPU: world_size-2 gloo, rendezvous on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from beacon_b200 import dist as bd


def test_shard_ranges_partition():
    for n, w in ((4096, 8), (1024, 3), (7, 2), (5, 8)):
        spans = [bd.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        bd.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = bd.shard_range(num_envs, rank, world)
        obs = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
        rwd = -torch.arange(lo, hi, dtype=torch.float64)
        g_obs = bd.gather_to_learner(obs, num_envs, dst=0)
        g_all = bd.all_gather_rows(rwd, num_envs)
        full = torch.arange(num_envs, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 10.0]], dtype=torch.float64) if rank == 0 else None
        mine = bd.scatter_actions(full, num_envs, src=0, like=torch.empty(0, 2, dtype=torch.float64))
        ok = torch.equal(g_all, -torch.arange(num_envs, dtype=torch.float64))
        ok &= torch.equal(mine[:, 0], torch.arange(lo, hi, dtype=torch.float64)) and torch.equal(mine[:, 1], 10 * torch.arange(lo, hi, dtype=torch.float64))
        if rank == 0:
            ok &= torch.equal(g_obs[:, 0], torch.arange(num_envs, dtype=torch.float64)) and g_obs.shape == (num_envs, 3)
        else:
            ok &= g_obs is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_envs", [8, 7])
def test_gather_scatter_world2_gloo(num_envs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_envs, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]
