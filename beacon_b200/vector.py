"""Callers of the hot path: a gymnasium-VectorEnv-like wrapper with auto-reset,
and the batched form of the separable (one policy per jet) protocol of shkadov_separable.

Everything stays on the device: observations / rewards / flags are CUDA tensors.
"""
import torch

from .batched import BatchedEnv


class VectorEnv:
    """B environments with auto-reset ("next-step" convention of gymnasium >= 1.0 is NOT used:
    like gymnasium 0.29, an env that finishes is reset inside the same step() call; the terminal
    observation is returned in info["final_obs"], and obs holds the first observation of the new
    episode).  shkadov envs restart with a random number of warm steps U{0..rand_steps}
    (shkadov.py:118-123) drawn from a seeded device generator."""

    def __init__(self, name, num_envs, auto_reset=True, rand_init=True, rand_steps=400, seed=0, device=0, **kwargs):
        self.env = BatchedEnv(name, batch=num_envs, seed=seed, device=device, **kwargs)
        self.name, self.num_envs, self.auto_reset = name, num_envs, auto_reset
        self.rand_init, self.rand_steps = rand_init and name == "shkadov", rand_steps
        self._gen = torch.Generator(device=self.env.device)
        self._gen.manual_seed(seed)
        self.episode_return = torch.zeros(num_envs, dtype=self.env.dtype, device=self.env.device)
        self.episode_length = torch.zeros(num_envs, dtype=torch.int64, device=self.env.device)

    @property
    def single_observation_dim(self):
        return self.env.n_obs

    def _n_warm(self):
        if not self.rand_init:
            return None
        return torch.randint(0, self.rand_steps + 1, (self.num_envs,), generator=self._gen, device=self.env.device,
                             dtype=torch.int32)

    def reset(self, mask=None, out=None):
        obs = self.env.reset(mask=mask, n_warm=self._n_warm(), max_warm=self.rand_steps if self.rand_init else None, out=out)
        if mask is None:
            self.episode_return.zero_(); self.episode_length.zero_()
        else:
            m = torch.as_tensor(mask, device=self.env.device).bool()
            self.episode_return.masked_fill_(m, 0); self.episode_length.masked_fill_(m, 0)
        return obs

    def step(self, actions, noise=None):
        """No device->host synchronisation: the masked reset is launched unconditionally after every step
        (its CTAs return at once for envs that are not done) and writes the first observation of the new
        episode over the rows of the finished envs; `info` holds device tensors — `final_obs`,
        `episode_return`, `episode_length` are meaningful where `done` is set."""
        obs, rwd, done, trunc = self.env.step(actions, noise=noise)
        r = rwd if rwd.dim() == 1 else rwd.sum(-1)
        self.episode_return += r
        self.episode_length += 1
        info = {}
        if self.auto_reset:
            info["final_obs"] = obs.clone()
            info["episode_return"] = self.episode_return.clone()
            info["episode_length"] = self.episode_length.clone()
            obs = self.reset(mask=done, out=obs)
        return obs, rwd, done, trunc, info

    def close(self):
        self.env.close()


class SeparableShkadov:
    """Batched shkadov_separable (shkadov.py:376-481): ONE solve per physical step, every jet is
    its own pseudo-environment with 10 observations and its own reward.

        reset() -> obs  [B, n_jets, n_obs]
        step(actions [B, n_jets]) -> obs [B, n_jets, n_obs], rwd [B, n_jets], done [B], trunc [B]

    `as_round_robin()` flattens to the reference's call order (jet-major per env): the k-th call of
    the reference's step() on env b returns `obs[b, k]`, `rwd[b, k]` of the same physical step."""

    def __init__(self, num_envs, n_jets=5, **kwargs):
        self.env = BatchedEnv("shkadov", batch=num_envs, n_jets=n_jets, per_jet_rwd=True, **kwargs)
        self.num_envs, self.n_jets = num_envs, n_jets
        self.n_obs = self.env.n_obs // n_jets

    def reset(self, mask=None, n_warm=None, out=None):
        """With a mask only the masked envs' rows are written (pass the current observations as `out`,
        [B, n_jets, n_obs] contiguous, to have them merged in place; otherwise the other rows are NaN)."""
        o = None if out is None else out.view(self.num_envs, self.n_jets * self.n_obs)
        return self.env.reset(mask=mask, n_warm=n_warm, out=o).view(self.num_envs, self.n_jets, self.n_obs)

    def step(self, actions, noise=None):
        obs, rwd, done, trunc = self.env.step(actions, noise=noise)
        return obs.view(self.num_envs, self.n_jets, self.n_obs), rwd.view(self.num_envs, self.n_jets), done, trunc

    @staticmethod
    def as_round_robin(obs, rwd):
        return obs.reshape(-1, obs.shape[-1]), rwd.reshape(-1)
