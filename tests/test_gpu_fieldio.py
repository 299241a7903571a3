"""Device-side init-state generation and the dump / load / warmup methods of the single-env
classes, against the oracle from the same rest state."""
import numpy as np
import pytest
import torch

from oracle import beacon_oracle as bo

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) / max(1.0, float(np.max(np.abs(b))))


def test_sloshing_init_generation_matches_oracle_from_rest():
    """sloshing/init.py:13-30: from h = 1, q = 0, n_warmup = 40 actions of the excitation signal."""
    from beacon_b200 import fieldio
    g = fieldio.generate_init_state("sloshing")
    o = bo.sloshing(init=False)
    o.reset()
    o.h[:] = 1.0
    o.q[:] = 0.0
    for it in range(40):
        o.step(np.array([fieldio.sloshing_signal(it * 0.05)]))
    assert rel(g["h"], o.h) <= 1e-10 and rel(g["q"], o.q) <= 1e-10
    assert abs(np.mean(g["h"][1:-1]) - 1.0) < 1e-2             # closed tank: volume is (nearly) conserved
    assert np.std(g["h"][1:-1]) > 0.05                          # and it sloshes


def test_shkadov_init_generation_from_rest_and_beyond_the_41_jet_cap(tmp_path):
    from beacon_b200 import fieldio
    from beacon_b200.envs import shkadov
    # short, noise-free warm-up against the oracle (shkadov/init.py:13-27 starts from reset_fields())
    g = fieldio.generate_init_state("shkadov", n_warmup=12, L0=150.0, n_jets=5, sigma=0.0)
    o = bo.shkadov(init=False, n_jets=5)
    o.reset()
    o.h[:] = 1.0
    o.q[:] = 1.0
    for _ in range(12):
        o.step(np.zeros(5))
    assert rel(g["h"], o.h) <= 1e-10 and rel(g["q"], o.q) <= 1e-10
    # the full 4000-action warm-up of a domain the shipped file cannot initialise (n_jets = 60, nx = 3850)
    g = fieldio.generate_init_state("shkadov", L0=150.0, n_jets=60, seed=5)
    assert g["h"].shape == (3850,) and np.all(np.isfinite(g["h"])) and np.all(np.isfinite(g["q"]))
    assert 0.9 < np.mean(g["h"]) < 1.1 and np.std(g["h"][2000:]) > 0.05      # developed waves downstream
    assert np.std(g["h"][:200]) < 0.02                                        # still flat near the inlet
    # write it in the reference's format, load it as the initial state of a 60-jet env and step
    p = tmp_path / "init_field.dat"
    fieldio.dump_fields("shkadov", p, **g)
    env = shkadov(init=False, n_jets=60)
    env.rand_init = False
    env.load(str(p))
    obs, _ = env.reset()
    assert obs.shape == (600,) and rel(env.h, np.loadtxt(p)[:, 1]) == 0.0
    obs, rwd, done, trunc, _ = env.step(np.zeros(60))
    assert np.isfinite(rwd) and rwd < 0.0 and env.status == 0


def test_rayleigh_warmup_and_dump_load_roundtrip(tmp_path):
    from beacon_b200 import fieldio
    from beacon_b200.envs import rayleigh
    g = fieldio.generate_init_state("rayleigh", n_warmup=2)     # rayleigh/init.py: from rest, n_sgts = 1, zero control
    o = bo.rayleigh(init=False, n_sgts=1)
    o.reset()
    for _ in range(2):
        o.step(np.zeros(1))
    for k in "uvpT":
        assert rel(g[k], getattr(o, k)) <= 1e-10, k
    # single-env class: warmup() repeats the previous action, dump() / load() use the reference's text format
    env = rayleigh()
    env.reset()
    env.step(np.linspace(-0.5, 0.5, 10))
    a_before = env.a.copy()
    env.warmup(3)
    o = bo.rayleigh()
    o.reset()
    for _ in range(4):
        o.step(np.linspace(-0.5, 0.5, 10))
    assert rel(env.T, o.T) <= 1e-10 and np.allclose(env.a, a_before, rtol=0, atol=1e-15)   # re-centred every action (rayleigh.py:165)
    f, fa = tmp_path / "f.dat", tmp_path / "a.dat"
    env.dump(str(f), str(fa))
    assert np.loadtxt(f).shape == (208, 52) and np.loadtxt(fa).shape == (10,)
    env2 = rayleigh()
    env2.load(str(f))
    env2.reset()
    assert rel(env2.T, env.T) < 5e-6 and rel(env2.u, env.u) < 5e-6      # %.5e precision
    assert np.array_equal(env2.T, np.loadtxt(f)[156:208])
