"""Generate tests/golden/<env>.npz from the UNMODIFIED reference (jviquerat/beacon).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/gen_golden.py            # all envs
    python oracle/gen_golden.py rayleigh   # one env

Each file holds the inputs (actions, injected noise, ctor kwargs) and the reference's outputs
(per-action full fields, observations, rewards, flags, Poisson sweep counts, derived integer
parameters).  The reference is driven through oracle/refload.py: its own classes, its own
numba kernels, float64 actions, numpy's global-RNG draws replaced by recorded numbers
(SURVEY.md Appendix B).  tests/test_oracle_golden.py replays these on the C oracle (bit-exact),
tests/test_*_gpu.py on the CUDA path (<= 1e-10).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import refload  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

SHK_INTS = ("nx", "ndt_act", "n_act", "n_interp", "jet_pos", "jet_hw", "jet_space", "jet_start",
            "jet_end", "l_rwd", "n_obs", "l_obs", "rwd_start", "rwd_end", "obs_start", "obs_end", "n_warmup")


def _save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def gen_shkadov():
    rng = np.random.default_rng(1001)
    out = {}
    # derived integer tables for several constructor argument sets (bit-exact contract)
    table = []
    for kw in ({}, {"n_jets": 10}, {"n_jets": 1}, {"n_jets": 20}, {"n_jets": 41},
               {"n_jets": 3, "jet_space": 7.0, "jet_pos": 120.0}, {"n_jets": 8, "L0": 200.0, "t_act": 10.0}):
        env, _ = refload.make("shkadov", **kw)
        table.append({"kwargs": kw, **{k: int(getattr(env, k)) for k in SHK_INTS}, "dx": float(env.dx)})
    out["params_json"] = np.array(json.dumps(table))

    for tag, n_jets, n_steps in (("j10", 10, 4), ("j5", 5, 2), ("j41", 41, 1)):
        env, mod = refload.make("shkadov", n_jets=n_jets)
        env.rand_init = False
        obs0, _ = env.reset()
        nd = env.ndt_act
        actions = rng.uniform(-1, 1, (n_steps, n_jets))
        noise = rng.uniform(-env.sigma, env.sigma, (n_steps, nd))
        H, Q, RH, RQ, OBS, RWD = [], [], [], [], [], []
        for k in range(n_steps):
            with refload.patched_noise(mod, noise[k]):
                obs, rwd, done, trunc, _ = env.step(actions[k].copy())
            H.append(env.h.copy()); Q.append(env.q.copy()); RH.append(env.rhsh.copy()); RQ.append(env.rhsq.copy())
            OBS.append(np.array(obs)); RWD.append(float(rwd))
        out.update({f"{tag}_actions": actions, f"{tag}_noise": noise, f"{tag}_obs0": obs0,
                    f"{tag}_h": np.array(H), f"{tag}_q": np.array(Q), f"{tag}_rhsh": np.array(RH),
                    f"{tag}_rhsq": np.array(RQ), f"{tag}_obs": np.array(OBS), f"{tag}_rwd": np.array(RWD)})

    # reset with warm steps: random.randint patched to a fixed count, noise recorded
    env, mod = refload.make("shkadov", n_jets=10)
    n_warm = 3
    noise = rng.uniform(-env.sigma, env.sigma, (n_warm, env.ndt_act))
    orig = mod.random.randint
    mod.random.randint = lambda a, b: n_warm
    try:
        with refload.patched_noise(mod, noise.reshape(-1)):
            obs0, _ = env.reset()
    finally:
        mod.random.randint = orig
    out.update({"warm_n": np.array(n_warm), "warm_noise": noise, "warm_obs0": obs0,
                "warm_h": env.h.copy(), "warm_q": env.q.copy(), "warm_stp": np.array(env.stp)})

    # KAT of SURVEY.md §4 (sigma=0, linspace actions)
    env, mod = refload.make("shkadov", n_jets=10)
    env.rand_init = False
    env.sigma = 0.0
    env.reset()
    sr = 0.0
    for f in (1.0, -0.5, 1.0):
        _, r, *_ = env.step(np.linspace(-1, 1, 10) * f)
        sr += r
    out.update({"kat_sum_rwd": np.array(sr), "kat_sum_h": np.array(env.h.sum()), "kat_sum_q": np.array(env.q.sum())})

    # horizon flags: t_act=0.2 -> n_act=4
    env, mod = refload.make("shkadov", n_jets=2, t_act=0.2)
    env.rand_init = False
    env.sigma = 0.0
    env.reset()
    flags = []
    for k in range(4):
        _, r, d, t, _ = env.step(np.zeros(2))
        flags.append((d, t))
    out["horizon_flags"] = np.array(flags)

    # separable protocol: n_jets=4, 2 physical steps => 8 calls (+4 reset calls)
    env, mod = refload.make("shkadov_separable", n_jets=4)
    env.rand_init = False
    env.sigma = 0.0
    r_obs = [env.reset()[0].copy() for _ in range(4)]
    acts = rng.uniform(-1, 1, (2, 4))
    s_obs, s_rwd, s_done = [], [], []
    for k in range(2):
        for j in range(4):
            o, r, d, t, _ = env.step(acts[k].copy())
            s_obs.append(o.copy()); s_rwd.append(r); s_done.append((d, t))
    out.update({"sep_reset_obs": np.array(r_obs), "sep_actions": acts, "sep_obs": np.array(s_obs),
                "sep_rwd": np.array(s_rwd), "sep_flags": np.array(s_done), "sep_h": env.h.copy(), "sep_q": env.q.copy()})
    _save("shkadov", **out)


def gen_burgers():
    rng = np.random.default_rng(1002)
    env, mod = refload.make("burgers")
    obs0, _ = env.reset()
    n = 6
    actions = rng.uniform(-1, 1, (n, 1))
    noise = rng.uniform(-env.sigma, env.sigma, n)
    U, UP, UPP, OBS, RWD = [], [], [], [], []
    for k in range(n):
        with refload.patched_noise(mod, noise[k:k + 1]):
            obs, rwd, *_ = env.step(actions[k].copy())
        U.append(env.u.copy()); UP.append(env.up.copy()); UPP.append(env.upp.copy()); OBS.append(obs.copy()); RWD.append(rwd)
    ints = {k: int(getattr(env, k)) for k in ("nx", "ctrl_pos", "ndt_act", "n_act", "n_obs_pts")}
    # KAT (sigma=0)
    env2, _ = refload.make("burgers", sigma=0.0)
    env2.reset()
    sr = 0.0
    for k in range(20):
        _, r, *_ = env2.step(np.array([0.3 if k % 2 == 0 else -0.6]))
        sr += r
    _save("burgers", actions=actions, noise=noise, obs0=obs0, u=np.array(U), up=np.array(UP), upp=np.array(UPP),
          obs=np.array(OBS), rwd=np.array(RWD), params_json=np.array(json.dumps({**ints, "dx": env.dx, "dt": env.dt})),
          kat_sum_rwd=np.array(sr), kat_sum_u=np.array(env2.u.sum()), kat_u=env2.u[250:255].copy())


def gen_sloshing():
    rng = np.random.default_rng(1003)
    env, mod = refload.make("sloshing")
    obs0, _ = env.reset()
    n = 6
    actions = rng.uniform(-1, 1, (n, 1))
    H, Q, RH, RQ, OBS, RWD = [], [], [], [], [], []
    for k in range(n):
        obs, rwd, *_ = env.step(actions[k].copy())
        H.append(env.h.copy()); Q.append(env.q.copy()); RH.append(env.rhsh.copy()); RQ.append(env.rhsq.copy())
        OBS.append(obs.copy()); RWD.append(rwd)
    ints = {k: int(getattr(env, k)) for k in ("nx", "ndt_act", "n_act", "n_interp", "n_obs")}
    env2, _ = refload.make("sloshing")
    env2.reset()
    sr = 0.0
    for k in range(5):
        _, r, *_ = env2.step(np.array([0.5 if k % 2 == 0 else -0.25]))
        sr += r
    _save("sloshing", actions=actions, obs0=obs0, h=np.array(H), q=np.array(Q), rhsh=np.array(RH), rhsq=np.array(RQ),
          obs=np.array(OBS), rwd=np.array(RWD), params_json=np.array(json.dumps({**ints, "dx": env.dx})),
          kat_sum_rwd=np.array(sr), kat_sum_h=np.array(env2.h.sum()), kat_sum_q=np.array(env2.q.sum()))


def gen_lorenz():
    rng = np.random.default_rng(1004)
    env, mod = refload.make("lorenz")
    obs0 = env.reset()[0].copy()
    n = 120
    actions = rng.integers(0, 3, n)
    X, FX, OBS, RWD, FL = [], [], [], [], []
    for k in range(n):
        obs, rwd, d, t, _ = env.step(np.int64(actions[k]))
        X.append(env.x.copy()); FX.append(env.fx.copy()); OBS.append(obs.copy()); RWD.append(rwd); FL.append((d, t))
    env2, _ = refload.make("lorenz")
    env2.reset()
    sr = 0.0
    for k in range(500):
        _, r, d, t, _ = env2.step(np.int64(k % 3))
        sr += r
    _save("lorenz", actions=actions, obs0=obs0, x=np.array(X), fx=np.array(FX), obs=np.array(OBS), rwd=np.array(RWD),
          kat500_sum_rwd=np.array(sr), kat500_last_flags=np.array((d, t)), n_act=np.array(env.n_act))


def gen_vortex():
    rng = np.random.default_rng(1005)
    env, mod = refload.make("vortex")
    obs0 = env.reset()[0].copy()
    n = 40
    actions = rng.uniform(-1, 1, (n, 2))
    X, FX, OBS, RWD = [], [], [], []
    for k in range(n):
        obs, rwd, *_ = env.step(actions[k].copy())
        X.append(env.x.copy()); FX.append(env.fx.copy()); OBS.append(obs.copy()); RWD.append(rwd)
    _save("vortex", actions=actions, obs0=obs0, x=np.array(X), fx=np.array(FX), obs=np.array(OBS), rwd=np.array(RWD),
          n_act=np.array(env.n_act), ndt_act=np.array(env.ndt_act))


def _run_mac(name, actions, int_keys):
    env, mod = refload.make(name)
    obs0 = env.reset()[0].copy()
    counts = []
    orig = mod.poisson

    def counting(*a):
        itp, ovf = orig(*a)
        counts.append(itp)
        return itp, ovf

    mod.poisson = counting
    scal = "T" if name == "rayleigh" else "C"
    F = {k: [] for k in ("u", "v", "p", scal)}
    OBS, RWD, ITP = [], [], []
    try:
        for a in actions:
            counts.clear()
            obs, rwd, *_ = env.step(np.array(a, dtype=np.float64).copy() if name == "rayleigh" else int(a))
            for k in F:
                F[k].append(getattr(env, k).copy())
            OBS.append(obs.copy()); RWD.append(rwd); ITP.append(np.array(counts))
    finally:
        mod.poisson = orig
    ints = {k: int(getattr(env, k)) for k in int_keys}
    arrs = {k: np.array(v) for k, v in F.items()}
    return dict(actions=np.array(actions), obs0=obs0, obs=np.array(OBS), rwd=np.array(RWD), itp=np.array(ITP),
                params_json=np.array(json.dumps(ints)), **arrs)


def gen_rayleigh():
    rng = np.random.default_rng(1006)
    acts = [np.zeros(10), np.linspace(-0.75, 0.75, 10), rng.uniform(-1, 1, 10)]
    _save("rayleigh", **_run_mac("rayleigh", acts, ("nx", "ny", "ndt_act", "n_act", "nx_sgts", "n_obs_tot", "nx_obs", "ny_obs")))


def gen_mixing():
    _save("mixing", **_run_mac("mixing", [0, 2], ("nx", "ny", "ndt_act", "n_act", "n_obs_tot", "nx_obs", "ny_obs")))


GEN = {"shkadov": gen_shkadov, "burgers": gen_burgers, "sloshing": gen_sloshing, "lorenz": gen_lorenz,
       "vortex": gen_vortex, "rayleigh": gen_rayleigh, "mixing": gen_mixing}

if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    for n in (sys.argv[1:] or list(GEN)):
        GEN[n]()
