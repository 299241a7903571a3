"""Worker of tests/test_gpu_dist.py (one process per GPU, launched by torch.distributed.run):
shards a shkadov batch over the ranks, steps it, and checks the three ways of bringing the rows to
the learner rank — NCCL gather, NCCL all-gather, and the fused peer-memory epilogue
(beacon_b200.peer.LearnerBuffer) — against the unsharded batch stepped on the learner's GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from beacon_b200 import BatchedEnv
    from beacon_b200 import dist as bd
    from beacon_b200.peer import LearnerBuffer

    for name, N, kw in (("shkadov", 7, dict(n_jets=5)), ("rayleigh", 4, dict())):
        env = bd.make_sharded(name, N, device=local, seed=5, **kw)
        env2 = bd.make_sharded(name, N, device=local, seed=5, **kw)     # same shard, stepped through the peer buffer
        lo, hi = bd.shard_range(N, rank, world)
        g = torch.Generator(device="cpu"); g.manual_seed(3)
        acts = torch.rand(3, N, env.act_dim, generator=g, dtype=torch.float64) * 2 - 1
        full = BatchedEnv(name, batch=N, device=local, seed=5, **kw) if rank == 0 else None
        o = env.reset(); env2.reset()
        go = bd.gather_to_learner(o, N, dst=0)
        if rank == 0:
            assert torch.equal(go, full.reset()), "reset obs"
        lb = LearnerBuffer(env2, N, dst=0)
        for k in range(3):
            mine = bd.scatter_actions(acts[k].to(dev) if rank == 0 else None, N, src=0, like=torch.empty(0, env.act_dim, dtype=torch.float64, device=dev))
            assert torch.equal(mine.cpu(), acts[k, lo:hi]), "scatter"
            obs, rwd, done, trunc = env.step(mine)
            g_obs = bd.gather_to_learner(obs, N, dst=0)
            g_rwd = bd.gather_to_learner(rwd, N, dst=0)
            a_done = bd.all_gather_rows(done.to(torch.uint8), N)
            lb.step(mine)
            lb.fence()
            if rank == 0:
                fo, fr, fd, ft = full.step(acts[k].to(dev))
                assert torch.equal(g_obs, fo) and torch.equal(g_rwd, fr), f"{name}: NCCL gather differs from the unsharded batch"
                assert torch.equal(a_done.bool(), fd)
                assert torch.equal(lb.obs, fo) and torch.equal(lb.rwd, fr), f"{name}: peer-memory epilogue differs"
                assert torch.equal(lb.done, fd) and torch.equal(lb.trunc, ft)
            else:
                assert g_obs is None and lb.obs is None
            dist.barrier()
        lb.close()
    if rank == 0:
        print("DIST_OK world", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
