"""CPU-side checks of the boundary: the C-ABI library builds/loads and exports every symbol the
header declares; host-side parameter derivation matches the reference's integers."""
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from beacon_b200 import _capi, build
    build.build()
    L = _capi.lib()
    hdr = open(os.path.join(ROOT, "include", "beacon_b200.h")).read()
    declared = set(re.findall(r"BEACON_API\s+[\w\s\*]+?\b(beacon_\w+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.beacon_version()


def test_struct_layouts_match_header():
    """ctypes mirrors must have the same field order as the C structs."""
    from beacon_b200 import _capi
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "beacon_b200.h")).read(), flags=re.S)
    for cname, st in (("beacon_shkadov_params", _capi.ShkadovParams), ("beacon_burgers_params", _capi.BurgersParams),
                      ("beacon_sloshing_params", _capi.SloshingParams), ("beacon_lorenz_params", _capi.LorenzParams),
                      ("beacon_vortex_params", _capi.VortexParams), ("beacon_mac_params", _capi.MacParams),
                      ("beacon_env_info_t", _capi.EnvInfo), ("beacon_common", _capi.Common)):
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname + ";", hdr).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(u?int\d+_t|double)\s+", "", decl)
            names += [re.sub(r"\[.*\]", "", n).strip() for n in decl.split(",")]
        assert names == [f[0] for f in st._fields_], cname


def test_no_cpu_fallback_without_gpu():
    import torch
    from beacon_b200 import BatchedEnv, BeaconError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(BeaconError):
        BatchedEnv("lorenz", batch=1)


def test_host_params_match_reference_integers(golden):
    from beacon_b200.params import BurgersCfg, LorenzCfg, MixingCfg, RayleighCfg, ShkadovCfg, SloshingCfg, VortexCfg
    g = golden("shkadov")
    for row in json.loads(str(g["params_json"])):
        c = ShkadovCfg(init=False, **row["kwargs"])
        for k in ("nx", "ndt_act", "n_act", "n_interp", "jet_pos", "jet_hw", "jet_space", "l_rwd", "n_obs", "l_obs", "n_warmup"):
            assert c.d[k] == row[k], (row["kwargs"], k)
        assert c.d["dx"] == row["dx"]
    P = json.loads(str(golden("burgers")["params_json"]))
    c = BurgersCfg()
    assert all(c.d[k] == P[k] for k in ("nx", "ctrl_pos", "ndt_act", "n_act", "n_obs_pts", "dx", "dt"))
    P = json.loads(str(golden("sloshing")["params_json"]))
    c = SloshingCfg()
    assert all(c.d[k] == P[k] for k in ("nx", "ndt_act", "n_act", "n_interp", "n_obs", "dx"))
    for name, cfg in (("rayleigh", RayleighCfg()), ("mixing", MixingCfg())):
        P = json.loads(str(golden(name)["params_json"]))
        assert all(cfg.d[k] == v for k, v in P.items()), name
    assert LorenzCfg().d["n_act"] == int(golden("lorenz")["n_act"])
    assert VortexCfg().d["n_act"] == int(golden("vortex")["n_act"]) and VortexCfg().d["ndt_act"] == int(golden("vortex")["ndt_act"])
    with pytest.raises(ValueError):
        ShkadovCfg(n_jets=42)      # shipped init file caps n_jets at 41


def test_philox_reference_vector():
    """Philox4x32-10 known-answer test (Random123 kat_vectors: counter=0,key=0 and the pi vector)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def philox(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        return c
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sass_census_identifies_the_hot_loops():
    """bench.py's in-run roofline multiplies these static per-loop counts by the run's trip counts: a wrong loop
    pick (e.g. a barrier loop that is not the Jacobi sweep) must not go unnoticed."""
    from beacon_b200 import build
    build.build()
    c = json.load(open(os.path.join(ROOT, "beacon_b200", "lib", "sass_census.json")))
    shk = c["shkadov_6_256_2"]["substep_unrolled2"]["per_substep"]
    assert 240 <= shk["fp64"] <= 320 and shk["instr"] < 700, shk
    for tag, cells in (("rayleigh_reg", 10), ("mixing_big", 20)):
        sw = c[tag]["per_sweep"]
        assert sw["bar"] == 1.0 and 6 * cells <= sw["fp64"] <= 9 * cells, (tag, sw)          # 6 fp64 per cell + the tile residual
        # (+ predicated: the big kernel's tile stores are predicated on "thread owns a tile")
        assert 4 * cells <= sw["smem_wavefronts"] + sw.get("smem_wavefronts_pred", 0) <= 8 * cells, (tag, sw)
        assert c[tag]["substep_other"]["fp64"] > 5 * sw["fp64"] and c[tag]["wavefront_loop"]["fp64"] >= 12
