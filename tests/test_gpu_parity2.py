"""Parity corners of round 2: fp32 builds of every env, whole-episode
returns and fields (SURVEY.md §8c protocol), fused launches of the large-grid kernel, OR-ed status bits,
checkpoint of the Philox draw counter, argument validation of the host-buffer step.  Everything goes
through the C-ABI; the oracle (oracle/, pinned bit-exact to the reference) is the checker."""
import numpy as np
import pytest
import torch

from oracle import beacon_oracle as bo
from test_gpu_parity import close, make, rep

pytestmark = pytest.mark.gpu

RTOL32 = 1e-5


# ----------------------------------------------------------------------------- fp32 builds (north star: <= 1e-5 per action)
def test_fp32_burgers_one_action(golden):
    g = golden("burgers")
    env = make("burgers", 2, dtype=torch.float32)
    env.reset()
    obs, rwd, d, t = env.step(rep(g["actions"][0], 2).float(), noise=rep(g["noise"][0:1], 2).float())
    close(env.get_state("u"), rep(g["u"][0], 2), rtol=RTOL32, what="burgers u fp32", floor=1.0)
    close(obs, rep(g["obs"][0], 2), rtol=RTOL32, what="burgers obs fp32", floor=1.0)
    assert abs(float(rwd[0]) - float(g["rwd"][0])) <= RTOL32 * max(1.0, abs(float(g["rwd"][0])))


def test_fp32_lorenz_vortex_one_action(golden):
    g = golden("lorenz")
    env = make("lorenz", 2, dtype=torch.float32)
    env.reset()
    obs, rwd, d, t = env.step(torch.full((2,), int(g["actions"][0])))
    close(obs, rep(g["obs"][0], 2), rtol=RTOL32, what="lorenz obs fp32", floor=1.0)
    close(env.get_state("x"), rep(g["x"][0], 2), rtol=RTOL32, what="lorenz x fp32", floor=1.0)
    g = golden("vortex")
    env = make("vortex", 2, dtype=torch.float32)
    env.reset()
    obs, rwd, d, t = env.step(rep(g["actions"][0], 2).float())
    close(env.get_state("x"), rep(g["x"][0], 2), rtol=RTOL32, what="vortex x fp32", floor=1.0)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name,scal,k", [("rayleigh", "T", 1), ("mixing", "C", 0)])
def test_fp32_mac2d_one_action(golden, name, scal, k):
    """fp32 arithmetic with the reference's tolerances (1e-8 / 1e-4 on the summed squared increment): the
    iteration must terminate well below itmax, and the fields after one multi-sweep action stay within
    1e-5 of the fp64 reference (the fp32 sweep counts differ by a few sweeps: not compared)."""
    g = golden(name)
    B = 2
    env = make(name, B, dtype=torch.float32)
    env.reset()
    if name == "rayleigh":
        env.step(rep(g["actions"][0], B).float())          # golden sequence: action 0 (zeros), then action 1
        act = rep(g["actions"][1], B).float()
    else:
        act = rep(g["actions"][0], B)
    obs, rwd, d, t = env.step(act, want_iters=True)
    assert int(env.status.max()) == 0, "fp32 Poisson iteration hit itmax"
    it, it_ref = int(env.last_iters[0, 0]), int(g["itp"][k].sum())
    assert abs(it - it_ref) <= 0.05 * it_ref, (it, it_ref)
    for f in ("u", "v", scal):
        close(env.get_state(f), rep(g[f][k].reshape(-1), B), rtol=RTOL32, what=f"{name} {f} fp32", floor=1.0)
    close(obs, rep(g["obs"][k], B), rtol=RTOL32, what="obs fp32", floor=1.0)


# ----------------------------------------------------------------------------- whole-episode returns and fields (§8c)
@pytest.mark.timeout(900)
def test_rayleigh_full_episode_return_and_fields():
    """100 free-running actions (a whole rayleigh-v0 episode) with random actions: sweep counts exact on every
    action, return within 1e-9, final fields within 1e-9 (rayleigh is not chaotic at Ra = 1e4)."""
    rng = np.random.default_rng(51)
    B, K = 2, 100
    env = make("rayleigh", B)
    orcs = [bo.rayleigh() for _ in range(B)]
    env.reset()
    [o.reset() for o in orcs]
    acts = rng.uniform(-1, 1, (K, B, 10))
    ret_g, ret_o = np.zeros(B), np.zeros(B)
    for k0 in range(0, K, 20):                               # 20 fused actions per launch
        obs, rwd, done, trunc = env.step_fused(torch.as_tensor(acts[k0:k0 + 20]), want_iters=True)
        it = env.last_iters.cpu().numpy()
        for k in range(20):
            ref = [o.step(acts[k0 + k, b].copy()) for b, o in enumerate(orcs)]
            assert [int(x) for x in it[k]] == [int(o.last_iters.sum()) for o in orcs], f"action {k0 + k}"
            ret_o += np.array([r[1] for r in ref])
            assert [bool(x) for x in done[k]] == [r[2] for r in ref]
        ret_g += rwd.sum(0).cpu().numpy()
    assert np.all(np.abs(ret_g - ret_o) <= 1e-9 * np.abs(ret_o)), (ret_g, ret_o)
    for f in ("u", "v", "p", "T"):
        close(env.get_state(f), np.stack([getattr(o, f).reshape(-1) for o in orcs]), rtol=1e-9, what=f"{f} after the episode")
    assert bool(done[-1].all()) and bool(trunc[-1].all())


@pytest.mark.timeout(900)
def test_mixing_20_actions_return_and_fused_launch():
    """20 free-running actions of mixing-v0 against the oracle (return within 1e-9, sweep counts exact), run as
    ten launches of K = 2 fused actions (the once-per-launch pressure-ghost path of the large-grid kernel);
    a second env doing the same 20 actions one launch each must agree bitwise except for p's ghost cells."""
    rng = np.random.default_rng(52)
    K = 20
    acts = rng.integers(0, 4, K)
    e1, e2 = make("mixing", 1), make("mixing", 1)
    o = bo.mixing()
    e1.reset(); e2.reset(); o.reset()
    ret_g = ret_o = 0.0
    for k0 in range(0, K, 2):
        a = torch.as_tensor(acts[k0:k0 + 2, None], dtype=torch.int32)
        obs, rwd, done, trunc = e1.step_fused(a, want_iters=True)
        it = e1.last_iters.cpu().numpy()
        for k in range(2):
            ro = o.step(int(acts[k0 + k]))
            o2, r2, _, _ = e2.step(a[k])
            assert int(it[k, 0]) == int(o.last_iters.sum()), f"action {k0 + k}"
            assert torch.equal(o2, obs[k]) and torch.equal(r2, rwd[k]), "K = 2 fused launch differs from two launches"
            ret_o += ro[1]
        ret_g += float(rwd.sum())
    assert abs(ret_g - ret_o) <= 1e-9 * abs(ret_o)
    for f in ("u", "v", "C"):
        assert torch.equal(e1.get_state(f), e2.get_state(f)), f
        close(e1.get_state(f)[0], getattr(o, f).reshape(-1), rtol=1e-9, what=f"mixing {f} after 20 actions")
    p1, p2 = e1.get_state("p").reshape(102, 102), e2.get_state("p").reshape(102, 102)
    assert torch.equal(p1[1:101, 1:101], p2[1:101, 1:101])
    close(p1.reshape(-1), o.p.reshape(-1), rtol=1e-9, what="mixing p after 20 actions")


def test_shkadov_full_episode_return():
    """sigma = 0, a whole 400-action episode free running: return within 1e-3 (chaotic film, §8c)."""
    rng = np.random.default_rng(53)
    env = make("shkadov", 1, n_jets=10, sigma=0.0)
    o = bo.shkadov(n_jets=10)
    env.reset(); o.reset()
    acts = rng.uniform(-1, 1, (400, 1, 10))
    tot_g = 0.0
    for k0 in range(0, 400, 100):
        _, r, done, trunc = env.step_fused(torch.as_tensor(acts[k0:k0 + 100]), noise=torch.zeros(100, 1, 50, dtype=torch.float64))
        tot_g += float(r.sum())
    tot_o = sum(o.step(acts[k, 0])[1] for k in range(400))
    assert abs(tot_g - tot_o) <= 1e-3 * abs(tot_o), (tot_g, tot_o)
    assert bool(done[-1, 0]) and bool(trunc[-1, 0]) and not bool(done[:-1].any())


# ----------------------------------------------------------------------------- status bits are OR-ed over the fused actions
def test_status_bits_or_over_fused_actions():
    """A blow-up flagged in the FIRST action of a K = 3 fused launch must survive the later actions that do
    not flag it (sloshing: a 60-cell bump of height 2.4 exceeds 2 h_max after action 1 only; the oracle
    gives done = True, False, False).  burgers: a NaN state flags NONFINITE."""
    o = bo.sloshing()
    o.reset()
    o.h[70:130] = 2.4
    flags = [o.step(np.zeros(1))[2] for _ in range(3)]
    assert flags == [True, False, False]
    env = make("sloshing", 2)
    env.reset()
    h = env.get_state("h")
    h[1, 70:130] = 2.4
    env.set_state("h", h)
    obs, rwd, done, trunc = env.step_fused(torch.zeros(3, 2, 1, dtype=torch.float64))
    assert [bool(x) for x in done[:, 1]] == flags and not bool(done[:, 0].any())
    assert int(env.status[1]) & 1, "BLOWUP of fused action 1 lost"
    assert int(env.status[0]) == 0
    close(env.get_state("h")[1], o.h, rtol=1e-9, what="h after the bump")
    env = make("burgers", 2)
    env.reset()
    u = env.get_state("u")
    u[0, 100] = float("nan")
    env.set_state("u", u)
    env.step_fused(torch.zeros(3, 2, 1, dtype=torch.float64), noise=torch.zeros(3, 2, 1, dtype=torch.float64))
    assert int(env.status[0]) & 4 and int(env.status[1]) == 0


# ----------------------------------------------------------------------------- ADVICE r1
@pytest.mark.parametrize("name,kw,adim", [("shkadov", dict(n_jets=5), 5), ("burgers", dict(), 1)])
def test_state_dict_carries_the_noise_stream(name, kw, adim):
    """step -> state_dict -> fresh env + load_state_dict -> step == uninterrupted run, with ON-DEVICE Philox
    noise: the per-env draw counter is part of the checkpoint."""
    acts = torch.rand(4, 3, adim, dtype=torch.float64, generator=torch.Generator().manual_seed(9)) * 2 - 1
    a = make(name, 3, seed=42, **kw)
    a.reset()
    a.step_fused(acts[:2])
    sd = a.state_dict()
    assert "draws" in sd and int(sd["draws"][0, 0]) > 0
    b = make(name, 3, seed=42, **kw)
    b.reset()
    b.load_state_dict(sd)
    oa, ob = a.step_fused(acts[2:]), b.step_fused(acts[2:])
    assert all(torch.equal(x, y) for x, y in zip(oa, ob))
    # the inlet noise has not reached the probes after two actions: compare the whole lattice
    fld = "h" if name == "shkadov" else "u"
    assert torch.equal(a.get_state(fld), b.get_state(fld)) and torch.equal(a.get_state("draws"), b.get_state("draws"))
    c = make(name, 3, seed=42, **kw)                        # without the counter the stream restarts: the inlet differs
    c.reset()
    c.load_state_dict({k: v for k, v in sd.items() if k != "draws"})
    c.step_fused(acts[2:])
    assert not torch.equal(c.get_state(fld), a.get_state(fld))


def test_step_host_validates_buffers():
    env = make("shkadov", 4, n_jets=5)
    env.reset()
    good = torch.zeros(4, 5, dtype=torch.float64)
    with pytest.raises(ValueError):
        env.step_host(torch.zeros(3, 5, dtype=torch.float64))                       # short action buffer
    with pytest.raises(ValueError):
        env.step_host(good, noise=torch.zeros(4, 49, dtype=torch.float64))          # short noise buffer
    out = list(env.alloc_host_outputs())
    out[0] = torch.empty(4, env.n_obs - 1, dtype=torch.float64)
    with pytest.raises(ValueError):
        env.step_host(good, out=tuple(out))                                         # short obs buffer
    out = list(env.alloc_host_outputs())
    out[1] = out[1].float()
    with pytest.raises(ValueError):
        env.step_host(good, out=tuple(out))                                         # wrong dtype
    with pytest.raises(ValueError):
        make("sloshing", 2).step_host(torch.zeros(2, 1, dtype=torch.float64), noise=torch.zeros(2, 1, dtype=torch.float64))
    o, r, d, t = env.step_host(good)
    assert d.dtype == torch.bool and t.dtype == torch.bool
    with pytest.raises(ValueError):
        make("lorenz", 1, dtype=torch.float16)


def test_masked_reset_rows_and_out_merge():
    env = make("sloshing", 4)
    o0 = env.reset()
    obs, _, _, _ = env.step(torch.full((4, 1), 0.5, dtype=torch.float64))
    mask = torch.tensor([True, False, True, False])
    r = env.reset(mask=mask)
    assert torch.equal(r[0], o0[0]) and torch.equal(r[2], o0[2]) and bool(torch.isnan(r[1]).all()) and bool(torch.isnan(r[3]).all())
    merged = env.reset(mask=mask, out=obs.clone())
    assert torch.equal(merged[0], o0[0]) and torch.equal(merged[1], obs[1]) and torch.equal(merged[3], obs[3])
    assert [int(x) for x in env.get_state("stp")[:, 0]] == [0, 1, 0, 1]


def test_single_env_warmup_keeps_step_counter():
    from beacon_b200.envs import mixing, sloshing
    e = sloshing()
    e.reset()
    e.step(np.array([0.3]))
    e.warmup(7)
    assert int(e._env.get_state("stp")[0, 0]) == 1 and e.stp == 1
    m = mixing()
    with pytest.raises(AttributeError):
        m.us                                                  # the register kernels do not keep the starred velocities


def test_shkadov_reset_longest_first_same_results():
    """The longest-first CTA order of a warm reset only changes the schedule: same state as the in-order launch."""
    nw = torch.tensor([3, 0, 5, 1, 5, 2, 0, 4], dtype=torch.int32)
    a = make("shkadov", 8, n_jets=5, seed=7)
    oa = a.reset(n_warm=nw)
    import subprocess, sys, os
    code = ("import sys; sys.path.insert(0, %r); import torch; from beacon_b200 import BatchedEnv; "
            "e = BatchedEnv('shkadov', batch=8, n_jets=5, seed=7); "
            "o = e.reset(n_warm=torch.tensor([3, 0, 5, 1, 5, 2, 0, 4], dtype=torch.int32)); "
            "torch.save((o.cpu(), e.get_state('h').cpu()), sys.argv[1])") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "r.pt")
        subprocess.run([sys.executable, "-c", code, f], check=True, env=dict(os.environ, BEACON_SHKADOV_RESET_INORDER="1"))
        ob, hb = torch.load(f)
    assert torch.equal(oa.cpu(), ob) and torch.equal(a.get_state("h").cpu(), hb)
