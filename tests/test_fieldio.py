"""Field files in the reference's text format: host logic, no GPU needed."""
import os

import numpy as np
import pytest

from beacon_b200 import fieldio
from beacon_b200.params import RayleighCfg, ShkadovCfg, SloshingCfg, init_fields


def test_shkadov_file_roundtrip_is_bit_identical(tmp_path):
    """the shipped init_field.dat values have 6 significant digits: dump -> load reproduces them
    exactly, and load() keeps the first nx rows like shkadov.py:364-368."""
    f = init_fields()
    c = ShkadovCfg(L0=550.0, n_jets=1)                       # the configuration init.py generates with: nx = 2900
    assert c.d["nx"] == f["shkadov_h"].shape[0] == 2900
    p = tmp_path / "init_field.dat"
    fieldio.dump_fields("shkadov", p, x=c.d["x"], h=f["shkadov_h"], q=f["shkadov_q"])
    g = fieldio.load_fields("shkadov", p)
    assert np.array_equal(g["h"], f["shkadov_h"]) and np.array_equal(g["q"], f["shkadov_q"])
    first = open(p).readline().split()
    assert len(first) == 3 and first[0] == "0.00000e+00" and first[1] == "%.5e" % f["shkadov_h"][0]
    g10 = fieldio.load_fields("shkadov", p, nx=1350)         # n_jets = 10 reads the first 1350 rows
    assert np.array_equal(g10["h"], ShkadovCfg(n_jets=10).d["h_init"])
    with pytest.raises(ValueError):
        fieldio.load_fields("shkadov", p, nx=3000)           # n_jets = 42 does not fit the shipped file


def test_rayleigh_and_sloshing_file_roundtrip(tmp_path):
    f = init_fields()
    p = tmp_path / "ray.dat"
    fieldio.dump_fields("rayleigh", p, u=f["rayleigh_u"], v=f["rayleigh_v"], p=f["rayleigh_p"], T=f["rayleigh_T"])
    a = np.loadtxt(p)
    assert a.shape == (208, 52)                              # 4 x 52 rows x 52 columns
    g = fieldio.load_fields("rayleigh", p)
    for k in "uvpT":
        assert np.array_equal(g[k], f["rayleigh_" + k]) and np.array_equal(g[k], RayleighCfg().d[k + "_init"])
    c = SloshingCfg()
    p = tmp_path / "slosh.dat"
    fieldio.dump_fields("sloshing", p, x=c.d["x"], h=c.d["h_init"], q=c.d["q_init"])
    assert np.loadtxt(p).shape == (200, 3)
    g = fieldio.load_fields("sloshing", p)
    assert np.array_equal(g["h"], c.d["h_init"]) and np.array_equal(g["q"], c.d["q_init"]) and g["h"][0] == 0.0
    with pytest.raises(ValueError):
        fieldio.dump_fields("lorenz", tmp_path / "x.dat")


def test_dump_precision_is_the_references(tmp_path):
    """%.5e keeps 6 significant digits (a dump is a lossy snapshot, not a checkpoint)."""
    rng = np.random.default_rng(0)
    u, v, pp, C = (rng.normal(size=(12, 12)) for _ in range(4))
    p = tmp_path / "mix.dat"
    fieldio.dump_fields("mixing", p, u=u, v=v, p=pp, C=C)
    g = fieldio.load_fields("mixing", p)
    assert np.max(np.abs(g["C"] - C) / np.abs(C)) < 5.1e-6 and not np.array_equal(g["C"], C)
    assert abs(fieldio.sloshing_signal(0.0) - 2.0) < 1e-15   # sloshing.py:134-138


def test_rest_states_and_warmup_lengths():
    """init.py starts from reset_fields(): shkadov h = q = 1 (shkadov.py:130-135), sloshing h = 1, q = 0
    (sloshing.py:100-104), rayleigh all zero (rayleigh.py:102-108); warm-up lengths int(t_warmup/dt_act)."""
    c = ShkadovCfg(init="rest", L0=550.0, n_jets=1)
    assert np.all(c.d["h_init"] == 1.0) and np.all(c.d["q_init"] == 1.0) and c.d["n_warmup"] == 4000
    s = SloshingCfg(init="rest")
    assert np.all(s.d["h_init"] == 1.0) and np.all(s.d["q_init"] == 0.0) and s.d["n_warmup"] == 40
    r = RayleighCfg(init=False, n_sgts=1)
    assert not r.d["T_init"].any() and r.d["n_warmup"] == 100 and r.d["nx_sgts"] == 50
    big = ShkadovCfg(init="rest", n_jets=60)                 # beyond the 41-jet cap of the shipped file
    assert big.d["nx"] == 3850 and big.d["h_init"].shape == (3850,)
