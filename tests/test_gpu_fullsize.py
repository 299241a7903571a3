"""BASELINE.json-size runs of the CUDA path: size-independent properties + anchors.

The oracle cannot step thousands of environments in seconds, so at the full batch sizes the
checks are the properties the domain offers (the envs of a batch are independent, so equal
inputs must give bitwise equal rows; a row must not depend on the batch it sits in; per-jet
rewards add up to the joint reward; stepping K actions in one launch equals K launches), each
tied to the reference by comparing a few rows with the committed golden vectors / the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import beacon_oracle as bo

pytestmark = pytest.mark.gpu


def make(name, batch, **kw):
    from beacon_b200 import BatchedEnv
    return BatchedEnv(name, batch=batch, **kw)


def relerr(a, b):
    a = np.asarray(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b))))


def test_shkadov_1024_envs_rows_independent_and_anchored(golden):
    """configs[1]: shkadov-v0, 10 jets, 1024 envs.  Every env gets its own action / inlet noise;
    rows 0, 511, 1023 are re-run alone (batch of 1) and must match bitwise; the rows that got
    the golden action / noise must reproduce the reference's fields."""
    g = golden("shkadov")
    B, nj = 1024, 10
    rng = np.random.default_rng(21)
    acts = rng.uniform(-1, 1, (B, nj))
    noise = rng.uniform(-5e-4, 5e-4, (B, 50))
    acts[7], noise[7] = g["j10_actions"][0], g["j10_noise"][0]
    acts[1000], noise[1000] = g["j10_actions"][0], g["j10_noise"][0]
    env = make("shkadov", B, n_jets=nj)
    env.reset()
    obs, rwd, done, trunc = env.step(torch.as_tensor(acts), noise=torch.as_tensor(noise))
    h, q = env.get_state("h"), env.get_state("q")
    assert int(env.status.max()) == 0 and not bool(done.any())
    assert torch.equal(h[7], h[1000]) and torch.equal(q[7], q[1000]) and torch.equal(obs[7], obs[1000])
    assert relerr(h[7], g["j10_h"][0]) <= 1e-10 and relerr(q[7], g["j10_q"][0]) <= 1e-10
    assert abs(float(rwd[7]) - float(g["j10_rwd"][0])) <= 1e-12
    for b in (0, 511, 1023):
        one = make("shkadov", 1, n_jets=nj)
        one.reset()
        o1, r1, _, _ = one.step(torch.as_tensor(acts[b:b + 1]), noise=torch.as_tensor(noise[b:b + 1]))
        assert torch.equal(one.get_state("h")[0], h[b]) and torch.equal(one.get_state("q")[0], q[b])
        assert torch.equal(o1[0], obs[b]) and torch.equal(r1[0], rwd[b])
    # mass-like sanity of every row: film thickness stays near 1, fields finite
    assert bool(torch.isfinite(h).all()) and float((h - 1).abs().max()) < 0.5


def test_shkadov_1024_envs_free_running_vs_oracle_sample():
    """four of 1024 envs followed by the oracle for 5 free-running actions (sigma = 0)."""
    B, nj, K = 1024, 10, 5
    rng = np.random.default_rng(22)
    acts = rng.uniform(-1, 1, (K, B, nj))
    env = make("shkadov", B, n_jets=nj, sigma=0.0)
    env.reset()
    obs, rwd, done, trunc = env.step_fused(torch.as_tensor(acts), noise=torch.zeros(K, B, 50, dtype=torch.float64))
    h = env.get_state("h")
    for b in (3, 400, 777, 1023):
        o = bo.shkadov(n_jets=nj)
        o.reset()
        ret = sum(o.step(acts[k, b])[1] for k in range(K))
        assert relerr(h[b], o.h) <= 1e-9, b          # 5 free-running actions: SURVEY.md sensitivity table
        assert abs(float(rwd[:, b].sum()) - ret) <= 1e-9 * abs(ret)


def test_separable_many_actuators_per_jet_rewards_add_up():
    """configs[2]: shkadov_separable-v0, 41 jets (nx = 2900), 512 envs = one GPU's share of 4096.
    The per-jet rewards of the separable env sum to the joint reward of the plain env from the
    same state, and both share fields bitwise."""
    B, nj = 512, 41
    rng = np.random.default_rng(23)
    acts = torch.as_tensor(rng.uniform(-1, 1, (2, B, nj)))
    noise = torch.as_tensor(rng.uniform(-5e-4, 5e-4, (2, B, 50)))
    sep = make("shkadov", B, n_jets=nj, per_jet_rwd=True)
    joint = make("shkadov", B, n_jets=nj)
    sep.reset(); joint.reset()
    o1, r1, d1, t1 = sep.step_fused(acts, noise)
    o2, r2, d2, t2 = joint.step_fused(acts, noise)
    assert r1.shape == (2, B, nj) and r2.shape[:2] == (2, B)
    assert torch.equal(sep.get_state("h"), joint.get_state("h")) and torch.equal(o1, o2)
    assert float((r1.sum(-1) - r2.reshape(2, B)).abs().max()) <= 1e-13
    assert int(sep.status.max()) == 0
    # anchor: one row against the oracle
    o = bo.shkadov(n_jets=nj)
    o.reset()
    for k in range(2):
        o.step(acts[k, 5].numpy(), noise=noise[k, 5].numpy())
    assert relerr(sep.get_state("h")[5], o.h) <= 1e-10


def test_rayleigh_4096_envs_two_actions(golden):
    """configs[4]: rayleigh-v0, 4096 envs.  Action 1 = the reference's first golden action on every env
    (zeros: one sweep per solve); action 2 = the SECOND golden action (11 993 Jacobi sweeps in the
    reference) on the even rows and independent random actions on the odd rows, five of which are
    followed by the oracle.  Equal actions -> bitwise equal rows and sweep counts, equal to the golden
    fields; the multi-sweep path at full batch is anchored on golden and oracle rows."""
    g = golden("rayleigh")
    B = 4096
    rng = np.random.default_rng(24)
    env = make("rayleigh", B)
    env.reset()
    a0 = np.broadcast_to(g["actions"][0], (B, 10)).copy()
    obs, rwd, done, trunc = env.step(torch.as_tensor(a0), want_iters=True)
    assert bool((env.last_iters[0] == int(g["itp"][0].sum())).all())
    acts = np.broadcast_to(g["actions"][1], (B, 10)).copy()
    odd = np.arange(1, B, 2)
    acts[odd] = rng.uniform(-1, 1, (odd.size, 10))
    obs, rwd, done, trunc = env.step(torch.as_tensor(acts), want_iters=True)
    it = env.last_iters[0]
    T = env.get_state("T")
    assert int(env.status.max()) == 0
    assert int(g["itp"][1].sum()) > 10000 and int(it[0]) == int(g["itp"][1].sum()), "golden action 1 sweep count"
    even = torch.arange(0, B, 2, device=T.device)
    assert bool((it[even] == it[0]).all())
    assert torch.equal(T[even], T[0:1].expand(even.numel(), -1)) and torch.equal(obs[even], obs[0:1].expand(even.numel(), -1))
    for f in ("u", "v", "p", "T"):
        assert relerr(env.get_state(f)[4094], g[f][1].reshape(-1)) <= 1e-10, f
    assert relerr(obs[2], g["obs"][1]) <= 1e-10
    assert abs(float(rwd[0]) - float(g["rwd"][1])) <= 1e-12 * abs(float(g["rwd"][1]))
    assert len(torch.unique(it[odd])) > 10            # data-dependent trip counts really differ across the batch
    for b in (1, 777, 2049, 3333, 4095):              # oracle rows with their own random action
        o = bo.rayleigh()
        o.reset()
        o.step(g["actions"][0].copy())
        ro = o.step(acts[b].copy())
        assert int(it[b]) == int(o.last_iters.sum()), f"row {b}: sweep count differs from the oracle"
        for f in ("u", "v", "p", "T"):
            assert relerr(env.get_state(f)[b], getattr(o, f).reshape(-1)) <= 1e-10, (b, f)
        assert relerr(obs[b], ro[0]) <= 1e-10 and abs(float(rwd[b]) - ro[1]) <= 1e-12 * abs(ro[1])


def test_mixing_1024_envs_one_action(golden):
    """configs[3]: mixing-v0, 1024 envs, the reference's first golden action on every env."""
    g = golden("mixing")
    B = 1024
    env = make("mixing", B)
    env.reset()
    a0 = int(g["actions"][0])
    obs, rwd, done, trunc = env.step(torch.full((B,), a0, dtype=torch.int32), want_iters=True)
    it = env.last_iters[0]
    C = env.get_state("C")
    assert int(env.status.max()) == 0
    assert bool((it == int(g["itp"][0].sum())).all()), "Jacobi sweep count differs from the reference"
    assert torch.equal(C, C[0:1].expand(B, -1))
    for f in ("u", "v", "p", "C"):
        assert relerr(env.get_state(f)[B - 1], g[f][0].reshape(-1)) <= 1e-10, f
    assert relerr(obs[17], g["obs"][0]) <= 1e-10
    assert abs(float(rwd[512]) - float(g["rwd"][0])) <= 1e-12


def test_rayleigh_fused_launch_equals_single_launches():
    """two actions in one launch == two launches (p ghost cells, which never feed back, are brought
    up to date once per launch: they may differ in the last bits, everything else is bitwise)."""
    rng = np.random.default_rng(25)
    B = 8
    acts = torch.as_tensor(rng.uniform(-1, 1, (2, B, 10)))
    e1, e2 = make("rayleigh", B), make("rayleigh", B)
    e1.reset(); e2.reset()
    o1, r1, _, _ = e1.step_fused(acts)
    outs = [e2.step(acts[k]) for k in range(2)]
    assert torch.equal(o1, torch.stack([o[0] for o in outs])) and torch.equal(r1, torch.stack([o[1] for o in outs]))
    for f in ("u", "v", "T"):
        assert torch.equal(e1.get_state(f), e2.get_state(f)), f
    p1, p2 = e1.get_state("p").reshape(B, 52, 52), e2.get_state("p").reshape(B, 52, 52)
    assert torch.equal(p1[:, 1:51, 1:51], p2[:, 1:51, 1:51])
    assert float((p1 - p2).abs().max()) <= 1e-12
