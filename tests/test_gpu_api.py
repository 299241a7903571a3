"""The reference-facing Python API on the GPU: single-env classes with the reference's names /
signatures / return tuples, the auto-reset vector wrapper and the separable protocol."""
import numpy as np
import pytest
import torch

from oracle import beacon_oracle as bo

pytestmark = pytest.mark.gpu


def test_single_env_classes_follow_the_gym_duck_type():
    from beacon_b200 import envs
    for name, act in (("shkadov", np.zeros(5)), ("burgers", np.array([0.1])), ("sloshing", np.array([0.2])),
                      ("lorenz", np.int64(2)), ("vortex", np.array([0.5, -0.25])), ("rayleigh", np.linspace(-1, 1, 10)),
                      ("mixing", 2)):
        env = envs.ENVS[name]()
        if name == "shkadov":
            env.rand_init = False
        obs, info = env.reset()
        assert info is None and isinstance(obs, np.ndarray) and obs.dtype == np.float64
        a0 = np.array(act, copy=True)
        out = env.step(act)
        assert len(out) == 5 and out[4] is None
        obs, rwd, done, trunc, _ = out
        assert isinstance(rwd, float) and isinstance(done, bool) and isinstance(trunc, bool)
        assert np.array_equal(np.asarray(act), a0), "caller's action must not be mutated"
        assert obs.shape == env.observation_space.shape and env.stp == 1
        env.close()


def test_shkadov_single_env_matches_oracle_kat():
    """SURVEY.md §4 KAT: n_jets=10, rand_init=False, sigma=0, 3 steps."""
    from beacon_b200.envs import shkadov
    env = shkadov(n_jets=10)
    env.rand_init = False
    env.sigma = 0.0
    env.reset()
    sr = 0.0
    for f in (1.0, -0.5, 1.0):
        sr += env.step(np.linspace(-1, 1, 10) * f)[1]
    assert abs(sr - (-0.016856704466031224)) < 1e-12
    assert abs(env.h.sum() - 1332.7062381888381) < 1e-8 and abs(env.q.sum() - 1346.4015915144832) < 1e-8
    assert env.h.shape == (1350,)


def test_separable_round_robin_matches_reference(golden):
    from beacon_b200.envs import shkadov_separable
    g = golden("shkadov")
    env = shkadov_separable(n_jets=4)
    env.rand_init = False
    env.sigma = 0.0
    r_obs = np.array([env.reset()[0] for _ in range(4)])
    assert np.max(np.abs(r_obs - g["sep_reset_obs"])) < 1e-12
    obs, rwd, flags = [], [], []
    for k in range(2):
        for j in range(4):
            o, r, d, t, _ = env.step(g["sep_actions"][k])
            obs.append(o); rwd.append(r); flags.append((d, t))
    assert np.max(np.abs(np.array(obs) - g["sep_obs"])) < 1e-10
    assert np.max(np.abs(np.array(rwd) - g["sep_rwd"])) < 1e-12
    assert np.array_equal(np.array(flags), g["sep_flags"]) and env.stp == 2 and env.count == 0


def test_rayleigh_mixing_single_env_kat():
    from beacon_b200.envs import mixing, rayleigh
    env = rayleigh()
    env.reset()
    r0 = env.step(np.zeros(10))[1]
    r1 = env.step(np.linspace(-0.75, 0.75, 10))[1]
    assert abs(r0 - (-2.162578835240081)) < 1e-10 and abs(r1 - (-3.1042715751735157)) < 1e-10
    assert abs(env.T.sum() - 74.51326908912912) < 1e-9 and env.T.shape == (52, 52)
    env = mixing()
    env.reset()
    r0 = env.step(0)[1]
    assert abs(r0 - (-0.3581561537496566)) < 1e-11


def test_vector_env_auto_reset():
    from beacon_b200.vector import VectorEnv
    v = VectorEnv("lorenz", 5)
    obs = v.reset()
    assert obs.shape == (5, 6)
    n = v.env.n_act
    for k in range(n):
        obs, rwd, done, trunc, info = v.step(torch.full((5,), k % 3, dtype=torch.int32))
        assert bool(done.all()) == (k == n - 1)
    assert "final_obs" in info and int(info["episode_length"][0]) == n
    assert torch.allclose(obs[:, :3].cpu(), torch.full((5, 3), 10.0, dtype=torch.float64))   # fresh episode
    assert int(v.env.get_state("stp").max()) == 0
    obs, rwd, done, trunc, info = v.step(torch.zeros(5, dtype=torch.int32))
    assert not bool(done.any()) and int(v.episode_length[0]) == 1

    # shkadov: short horizon, random warm restarts differ per env
    v = VectorEnv("shkadov", 4, n_jets=2, t_act=0.1, seed=3, rand_steps=3)
    v.reset()
    for k in range(2):
        obs, rwd, done, trunc, info = v.step(torch.zeros(4, 2, dtype=torch.float64))
    assert bool(done.all()) and bool(trunc.all()) and "final_obs" in info
    assert int(v.env.get_state("stp").max()) == 0


def test_separable_batched_wrapper():
    from beacon_b200.vector import SeparableShkadov
    s = SeparableShkadov(3, n_jets=4, sigma=0.0)
    obs = s.reset()
    assert obs.shape == (3, 4, 10)
    o = bo.shkadov_separable(n_jets=4)
    [o.reset() for _ in range(4)]
    a = np.linspace(-1, 1, 4)
    obs, rwd, done, trunc = s.step(torch.as_tensor(np.tile(a, (3, 1))), noise=torch.zeros(3, 50, dtype=torch.float64))
    ref = [o.step(a) for _ in range(4)]
    fo, fr = s.as_round_robin(obs, rwd)
    assert fo.shape == (12, 10) and fr.shape == (12,)
    assert np.max(np.abs(obs[1].cpu().numpy() - np.array([r[0] for r in ref]))) < 1e-10
    assert np.max(np.abs(rwd[2].cpu().numpy() - np.array([r[1] for r in ref]))) < 1e-12


def test_state_dict_roundtrip_resumes_bitwise():
    from beacon_b200 import BatchedEnv
    a = BatchedEnv("sloshing", batch=3)
    a.reset()
    acts = torch.rand(4, 3, 1, dtype=torch.float64, device="cuda") * 2 - 1
    a.step_fused(acts[:2])
    sd = a.state_dict()
    b = BatchedEnv("sloshing", batch=3)
    b.reset()
    b.load_state_dict(sd)
    oa = a.step_fused(acts[2:])
    ob = b.step_fused(acts[2:])
    assert all(torch.equal(x, y) for x, y in zip(oa, ob))
