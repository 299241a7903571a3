"""Learner-rank buffers in peer memory: the fused form of the only exchange of the path.

The north star's one collective is the gather of observations / rewards / flags of every rank's env
shard to the learner rank (the reference outputs that travel are get_obs / get_rwd,
shkadov.py:239-264, rayleigh.py:243-275).  `dist.gather_to_learner` does it with NCCL after the step.
Here the step kernel does it itself: the learner rank owns ONE device allocation holding
`obs [N, n_obs] | rwd [N, rwd_dim] | done [N] | trunc [N]` for all N envs of the job, exports it over
CUDA IPC, every other process maps it (NVLink peer access on one node), and each rank's
`beacon_env_step` gets obs / rwd / done / trunc pointers INSIDE the mapping, offset to its own env
slice.  Every env's CTA then writes its rows into the learner GPU's HBM when it finishes —
overlapped with the compute of the remaining envs, exactly like the zero-copy host path of
`step_host` — and no gather follows; the only synchronisation is a stream-ordered barrier
(`LearnerBuffer.fence()`), after which the learner reads plain local tensors.

All device memory comes from the C-ABI (`beacon_peer_alloc` / `_export` / `_open`): torch is used for
the rendezvous (`broadcast_object_list` of the 64-byte handle) and for tensor views of the learner's
own allocation.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi as capi
from .dist import shard_range


class _DevMem:
    """__cuda_array_interface__ shim: lets torch wrap a raw device allocation of the C-ABI."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class LearnerBuffer:
    """`LearnerBuffer(env, num_envs, dst=0)` — collective over the default process group.

    env       this rank's BatchedEnv (its shard: `dist.shard_range(num_envs, rank, world)` rows)
    step(actions, noise=None)   one gym step of the shard, rows written into the learner's buffer
    fence()   stream-ordered completion: after it, on `dst`, `.obs/.rwd/.done/.trunc` hold the rows of
              ALL ranks for this step (torch views of the learner's allocation; None on other ranks)
    """

    def __init__(self, env, num_envs, dst=0, group=None):
        self.env, self.num_envs, self.dst, self.group = env, int(num_envs), dst, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lo, hi = shard_range(self.num_envs, self.rank, self.world)
        if hi - lo != env.batch:
            raise ValueError(f"rank {self.rank}: env holds {env.batch} envs, its shard of {num_envs} is {hi - lo}")
        self.lo = lo
        rb = 8 if env.dtype == torch.float64 else 4
        N = self.num_envs
        align = lambda x: (x + 255) // 256 * 256
        self.off_obs = 0
        self.off_rwd = align(N * env.n_obs * rb)
        self.off_done = self.off_rwd + align(N * env.rwd_dim * rb)
        self.off_trunc = self.off_done + align(N)
        self.nbytes = self.off_trunc + align(N)
        self._rb = rb
        self._lib = capi.lib()
        self._dev = env.device.index
        p = C.c_void_p()
        handle = [None]
        self._owned = self.rank == dst
        if self._owned:
            capi.check(self._lib.beacon_peer_alloc(self._dev, self.nbytes, C.byref(p)))
            buf = C.create_string_buffer(capi.PEER_HANDLE_BYTES)
            capi.check(self._lib.beacon_peer_export(p, buf))
            handle[0] = buf.raw
        dist.broadcast_object_list(handle, src=dst, group=group)
        if not self._owned:
            capi.check(self._lib.beacon_peer_open(handle[0], self._dev, C.byref(p)))
        self._base = int(p.value)
        self._sig = torch.zeros(1, dtype=torch.int32, device=env.device)
        self.obs = self.rwd = self.done = self.trunc = None
        if self._owned:
            raw = torch.as_tensor(_DevMem(self._base, self.nbytes), device=env.device)
            self._raw = raw
            self.obs = raw[self.off_obs:self.off_obs + N * env.n_obs * rb].view(env.dtype).view(N, env.n_obs)
            rwd = raw[self.off_rwd:self.off_rwd + N * env.rwd_dim * rb].view(env.dtype).view(N, env.rwd_dim)
            self.rwd = rwd[:, 0] if env.rwd_dim == 1 else rwd
            self.done = raw[self.off_done:self.off_done + N].view(torch.bool)
            self.trunc = raw[self.off_trunc:self.off_trunc + N].view(torch.bool)

    def step(self, actions, noise=None):
        e, lo, rb = self.env, self.lo, self._rb
        e.step_into(actions, self._base + self.off_obs + lo * e.n_obs * rb, self._base + self.off_rwd + lo * e.rwd_dim * rb,
                    self._base + self.off_done + lo, self._base + self.off_trunc + lo, noise=noise)

    def fence(self):
        """A 4-byte all-reduce on the stepping stream: it starts on a rank when that rank's step kernel
        (and with it every peer write of its epilogues) has completed, and completes on the learner
        only after every rank has joined — the rows of this step are then visible to the learner's
        following kernels.  Also keeps fast ranks from overwriting rows the learner still reads."""
        dist.all_reduce(self._sig, group=self.group)

    def close(self):
        if self._base:
            if self._owned:
                self.obs = self.rwd = self.done = self.trunc = self._raw = None
                torch.cuda.synchronize(self.env.device)
                dist.barrier(group=self.group)          # nobody writes any more
                self._lib.beacon_peer_free(self._dev, C.c_void_p(self._base))
            else:
                torch.cuda.synchronize(self.env.device)
                self._lib.beacon_peer_close(C.c_void_p(self._base))
                dist.barrier(group=self.group)
            self._base = 0
