"""Host-side derivation of every env's parameters from the reference's constructor arguments.

All derived integers (lattice indices, sub-step counts, probe positions) are computed HERE, in
Python, with the same float expressions and `int()` truncations the reference uses, and are then
passed as ints across the C-ABI: probe indices and actuator masks are bit-exact by construction
(SURVEY.md §5 "Config / flags").  Citations are file:line under /root/reference/beacon/.
"""
import math
import os
from dataclasses import dataclass, field

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "init_fields.npz")
_init_cache = None


def init_fields():
    """Developed-flow initial states shipped by the reference as init_field.dat (parsed once by
    tools/import_init_fields.py with np.loadtxt, exactly like the reference's load())."""
    global _init_cache
    if _init_cache is None:
        with np.load(_DATA) as f:
            _init_cache = {k: f[k].copy() for k in f.files}
    return _init_cache


def load_init_file(env, filename):
    """Reference text format ('%.5e'): shkadov/sloshing columns (x,h,q) — shkadov.py:364-368,
    sloshing.py:310-314; rayleigh u,v,p,T stacked vertically — rayleigh.py:356-362."""
    f = np.loadtxt(filename)
    if env in ("shkadov", "sloshing"):
        return {"h": f[:, 1].copy(), "q": f[:, 2].copy()}
    if env == "rayleigh":
        n = f.shape[0] // 4
        return {"u": f[0:n].copy(), "v": f[n:2 * n].copy(), "p": f[2 * n:3 * n].copy(), "T": f[3 * n:4 * n].copy()}
    raise ValueError(env)


@dataclass
class ShkadovCfg:
    """shkadov.py:20-76."""
    init: object = True          # True: shipped init_field.dat, False: zeros (as the reference), "rest": h = q = 1
    L0: float = 150.0
    n_jets: int = 5
    jet_pos: float = 150.0
    jet_space: float = 10.0
    delta: float = 0.1
    t_act: float = 20.0
    sigma: float = 5.0e-4
    per_jet_rwd: bool = False
    d: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        d = self.d
        L = self.L0 + self.jet_space * (self.n_jets + 2)            # :32
        nx = int(5 * L)                                             # :33
        dt, dt_act, t_warmup = 0.001, 0.05, 200.0
        jet_hw, l_obs, l_rwd, u_interp = 2.0, 10.0, 10.0, 0.02
        dx = float(L / nx)                                          # :57
        d.update(L=L, nx=nx, dx=dx, dt=dt, eps=1.0e-8, jet_amp=5.0, h_max=5.0, blowup_rwd=-1.0, rand_steps=400,
                 ndt_act=int(dt_act / dt), n_act=int(self.t_act / dt_act), n_warmup=int(t_warmup / dt_act),
                 n_interp=int(u_interp / dt), jet_pos=int(self.jet_pos / dx), jet_hw=int(jet_hw / dx),
                 jet_space=int(self.jet_space / dx), l_rwd=int(l_rwd / dx), n_obs=int(l_obs),
                 l_obs=int(l_obs / dx), obs_stride=int(1.0 / dx))   # :59-76, :246
        d["x"] = np.linspace(0, nx, num=nx, endpoint=False) * dx    # :79
        if isinstance(self.init, str) and self.init == "rest":      # reset_fields() state, :130-135 (init.py starts here)
            d["h_init"], d["q_init"] = np.ones(nx), np.ones(nx)
        elif self.init:
            f = init_fields()
            if nx > f["shkadov_h"].shape[0]:                        # same failure as load(), :366
                raise ValueError(f"init field has {f['shkadov_h'].shape[0]} points, nx={nx} needs more "
                                 "(n_jets <= 41 with the shipped file)")
            d["h_init"], d["q_init"] = f["shkadov_h"][:nx].copy(), f["shkadov_q"][:nx].copy()
        else:
            d["h_init"], d["q_init"] = np.zeros(nx), np.zeros(nx)

    def jet_mask(self):
        """Integer actuator table: (jet index, first lattice point, last lattice point)."""
        d = self.d
        return [(j, d["jet_pos"] + j * d["jet_space"] - d["jet_hw"], d["jet_pos"] + j * d["jet_space"] + d["jet_hw"])
                for j in range(self.n_jets)]

    def obs_indices(self):
        d = self.d
        return np.array([[d["jet_pos"] + j * d["jet_space"] - d["l_obs"] + k * d["obs_stride"]
                          for k in range(d["n_obs"])] for j in range(self.n_jets)], dtype=np.int64)


@dataclass
class BurgersCfg:
    """burgers.py:21-43."""
    u_target: float = 0.5
    amp: float = 10.0
    sigma: float = 0.1
    ctrl_pos: float = 1.0
    L: float = 2.0
    d: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        nx, t_max, dt_act, n_obs_pts = 500, 10.0, 0.05, 5
        dx = float(self.L / nx)
        dt = 0.2 * dx
        self.d.update(nx=nx, dx=dx, dt=dt, ctrl_pos=int(self.ctrl_pos / dx), ndt_act=int(dt_act / dt),
                      n_act=int(t_max / dt_act), n_obs_pts=n_obs_pts)


@dataclass
class SloshingCfg:
    """sloshing.py:19-50."""
    init: bool = True
    L: float = 2.5
    amp: float = 5.0
    alpha: float = 0.0005
    g: float = 9.81
    d: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        nx = int(80 * self.L)
        dt, dt_act, t_act, u_interp, obs_smpl = 0.001, 0.05, 10.0, 0.01, 2
        self.d.update(nx=nx, dx=float(self.L / nx), dt=dt, ndt_act=int(dt_act / dt), n_act=int(t_act / dt_act),
                      n_interp=int(u_interp / dt), obs_smpl=obs_smpl,
                      n_obs=nx // obs_smpl + (1 if (nx % obs_smpl != 0) else 0), h_max=1.0)
        self.d.update(dt_act=dt_act, n_warmup=int(2.0 / dt_act),                                   # sloshing.py:27,47
                      x=np.linspace(0, nx, num=nx, endpoint=False) * float(self.L / nx))           # :51
        h0, q0 = np.zeros(nx + 2), np.zeros(nx + 2)     # ghosts stay 0 until the first BC (sloshing.py:64-65)
        if isinstance(self.init, str) and self.init == "rest":                                     # reset_fields(), :100-104
            h0[:] = 1.0
        elif self.init:
            f = init_fields()
            if f["sloshing_h"].shape[0] != nx:
                raise ValueError("shipped sloshing init field has 200 cells (L=2.5)")
            h0[1:nx + 1], q0[1:nx + 1] = f["sloshing_h"], f["sloshing_q"]
        self.d["h_init"], self.d["q_init"] = h0, q0


@dataclass
class LorenzCfg:
    """lorenz.py:22-41."""
    sigma: float = 10.0
    rho: float = 28.0
    beta: float = 8.0 / 3.0
    d: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        dt, dt_act, t_max = 0.05, 0.05, 25.0
        self.d.update(dt=dt, ndt_act=int(dt_act / dt), n_act=int(t_max / dt_act), x0=(10.0, 10.0, 10.0),
                      forcing=(-1.0, 0.0, 1.0))


@dataclass
class VortexCfg:
    """vortex.py:21-61."""
    re: float = 50.0
    weight: float = 50.0
    d: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        omega_s, omega_f, beta, mass = 1.1, 0.74, 1.0, 10.0
        dt, dt_act, t_max = 0.1, 0.5, 400.0
        self.d.update(lmbda_re=9.153, lmbda_cx=3.239, mu_re=308.9, mu_cx=-1025.0, alpha_re=0.03492, alpha_cx=0.01472,
                      ire=1.0 / 46.6 - 1.0 / self.re, omega_s=omega_s, omega_f=omega_f, domega=omega_s - omega_f,
                      gamma=0.023, beta_m=beta / (omega_f * mass), weight=self.weight, dt=dt,
                      ndt_act=int(dt_act / dt), n_act=int(t_max / dt_act), mod_min=0.0, mod_max=0.3,
                      phase_min=-math.pi, phase_max=math.pi, x0=(-0.00385, -0.00378, 0.00118, -0.00131))


@dataclass
class RayleighCfg:
    """rayleigh.py:20-56."""
    init: bool = True
    L: float = 1.0
    H: float = 1.0
    n_sgts: int = 10
    ra: float = 1.0e4
    d: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        nx, ny = int(50 * self.L), int(50 * self.H)
        dt, dt_act, t_act = 0.01, 2.0, 200.0
        nxp, nyp = 4 * int(self.L), 4 * int(self.H)
        self.d.update(nx=nx, ny=ny, dx=float(self.L / nx), dy=float(self.H / ny), dt=dt, pr=0.71, ra=self.ra,
                      Tc=-0.5, Th=0.5, C=0.75, ndt_act=int(dt_act / dt), n_act=int(t_act / dt_act),
                      nx_sgts=nx // self.n_sgts, nx_obs_pts=nxp, ny_obs_pts=nyp, n_obs_steps=4,
                      nx_obs=nx // nxp, ny_obs=ny // nyp, n_obs_tot=3 * 4 * nxp * nyp, tol=1.0e-8, itmax=300000,
                      n_warmup=int(200.0 / dt_act))                                                # rayleigh.py:35,51
        shape = (nx + 2, ny + 2)
        if self.init:
            f = init_fields()
            if f["rayleigh_u"].shape != shape:
                raise ValueError("shipped rayleigh init field is 52x52 (L=H=1)")
            for k in "uvpT":
                self.d[k + "_init"] = f["rayleigh_" + k].copy()
        else:
            for k in "uvpT":
                self.d[k + "_init"] = np.zeros(shape)

    def probe_indices(self):
        d = self.d
        return [(d["nx_obs"] // 2 + i * d["nx_obs"], d["ny_obs"] // 2 + j * d["ny_obs"])
                for i in range(d["nx_obs_pts"]) for j in range(d["ny_obs_pts"])]


@dataclass
class MixingCfg:
    """mixing.py:21-49, reset_fields :82-111."""
    L: float = 1.0
    H: float = 1.0
    re: float = 100.0
    pe: float = 10000.0
    side: float = 0.5
    C0: float = 1.0
    d: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        nx, ny = int(100 * self.L), int(100 * self.H)
        dt, dt_act, t_act, nu = 0.002, 0.5, 50.0, 0.01
        nxp, nyp = 4 * int(self.L), 4 * int(self.H)
        dx, dy = float(self.L / nx), float(self.H / ny)
        self.d.update(nx=nx, ny=ny, dx=dx, dy=dy, dt=dt, re=self.re, pe=self.pe, u_max=self.re * nu / self.L,
                      ndt_act=int(dt_act / dt), n_act=int(t_act / dt_act), nx_obs_pts=nxp, ny_obs_pts=nyp,
                      n_obs_steps=4, nx_obs=nx // nxp, ny_obs=ny // nyp, n_obs_tot=3 * 4 * nxp * nyp,
                      tol=1.0e-4, itmax=300000, ref_c=(self.side * self.side) / (self.L * self.H) * self.C0)
        i_min = math.floor(0.5 * (self.L - self.side) / dx)            # :90-94 (ghost-inclusive indices)
        i_max = i_min + math.floor(self.side / dx)
        j_min = math.floor(0.5 * (self.H - self.side) / dy)
        j_max = j_min + math.floor(self.side / dy)
        C = np.zeros((nx + 2, ny + 2))
        C[i_min:i_max, j_min:j_max] = self.C0
        self.d.update(C_init=C, patch=(i_min, i_max, j_min, j_max))

    def probe_indices(self):
        d = self.d
        return [(d["nx_obs"] // 2 + i * d["nx_obs"], d["ny_obs"] // 2 + j * d["ny_obs"])
                for i in range(d["nx_obs_pts"]) for j in range(d["ny_obs_pts"])]


CFG = {"shkadov": ShkadovCfg, "burgers": BurgersCfg, "sloshing": SloshingCfg, "lorenz": LorenzCfg,
       "vortex": VortexCfg, "rayleigh": RayleighCfg, "mixing": MixingCfg}
