"""On-disk formats either side of the hot path, and device-side init-state generation
(SURVEY.md §8f rows 3 and 4).

* `dump_fields` / `load_fields`: the reference's text format for fields, `np.savetxt(fmt='%.5e')`
  — shkadov.py:353-368 (columns x, h, q), sloshing.py:298-314 (columns x, h[1:-1], q[1:-1]),
  rayleigh.py:344-362 and mixing.py:362-373 (u, v, p, T|C stacked vertically, (nx+2) rows each).
  Files written here can be read by the reference's `load()` and rendered with its own tools;
  the shipped `init_field.dat` files read back bit-identically (tests/test_fieldio.py).
* `generate_init_state`: what the reference's `init.py` scripts do on one CPU core in minutes
  (shkadov/init.py:13-27, rayleigh/init.py:13-28, sloshing/init.py:13-30) — start from rest, run
  `n_warmup` uncontrolled (or, for sloshing, excited) actions, keep the developed fields — done on
  the GPU through the same kernels as `step()`.  This lifts the n_jets <= 41 cap of the shipped
  shkadov file (nx <= 2900): any `L0` / `n_jets` can be warmed up in about a second.
The binary checkpoint of the full device state is `BatchedEnv.state_dict()` / `load_state_dict()`.
"""
import numpy as np

FMT = "%.5e"


def dump_fields(env, path, **f):
    """Write `path` in the reference's text format for `env`.

    shkadov: x, h, q        sloshing: x, h, q (h, q with their two ghost cells)
    rayleigh: u, v, p, T    mixing: u, v, p, C          (2D arrays of shape (nx+2, ny+2))"""
    if env in ("shkadov", "shkadov_separable"):
        arr = np.transpose(np.vstack((f["x"], f["h"], f["q"])))                      # shkadov.py:355-360
    elif env == "sloshing":
        h, q = np.asarray(f["h"]), np.asarray(f["q"])
        arr = np.transpose(np.vstack((f["x"], h[1:-1], q[1:-1])))                    # sloshing.py:300-305
    elif env in ("rayleigh", "mixing"):
        s = f["T"] if env == "rayleigh" else f["C"]
        arr = np.vstack((f["u"], f["v"], f["p"], s))                                 # rayleigh.py:346-351
    else:
        raise ValueError(f"{env}: the reference defines no field dump for this env")
    np.savetxt(path, arr, fmt=FMT)


def load_fields(env, path, nx=None):
    """Read a reference-format field file; returns the arrays `load()` would install as the
    initial state (shkadov.py:364-368, sloshing.py:310-314, rayleigh.py:356-362)."""
    a = np.loadtxt(path)
    if env in ("shkadov", "shkadov_separable"):
        n = a.shape[0] if nx is None else nx
        if a.shape[0] < n:
            raise ValueError(f"{path} has {a.shape[0]} points, nx={n} needs more")
        return {"x": a[:n, 0].copy(), "h": a[:n, 1].copy(), "q": a[:n, 2].copy()}
    if env == "sloshing":
        n = a.shape[0]
        h, q = np.zeros(n + 2), np.zeros(n + 2)                                      # ghosts stay 0 until the first BC
        h[1:-1], q[1:-1] = a[:, 1], a[:, 2]
        return {"x": a[:, 0].copy(), "h": h, "q": q}
    if env in ("rayleigh", "mixing"):
        if a.shape[0] % 4:
            raise ValueError(f"{path}: expected 4 stacked fields, got {a.shape[0]} rows")
        m = a.shape[0] // 4
        names = ("u", "v", "p", "T" if env == "rayleigh" else "C")
        return {k: a[i * m:(i + 1) * m].copy() for i, k in enumerate(names)}
    raise ValueError(f"{env}: the reference defines no field file for this env")


def sloshing_signal(t):
    """excitation of the sloshing init phase, sloshing.py:134-138."""
    return 0.5 * (np.cos(np.pi * t) + 3.0 * np.cos(4.0 * np.pi * t))


def generate_init_state(env, device=0, seed=0, n_warmup=None, chunk=50, **kwargs):
    """Developed-flow initial state from rest, on the GPU (the reference's init.py scripts).

    shkadov:  shkadov(init=False, L0=550, n_jets=1) from h = q = 1, n_warmup zero-control actions
              with inlet noise                                   -> dict(x, h, q)
    rayleigh: rayleigh(init=False, n_sgts=1) from rest, n_warmup zero-control actions -> dict(u, v, p, T)
    sloshing: sloshing(init=False) from h = 1, q = 0, n_warmup actions of `sloshing_signal` -> dict(x, h, q)
    `kwargs` are the env ctor's (e.g. L0=..., n_jets=... to warm up a longer shkadov domain)."""
    import torch
    from .batched import BatchedEnv

    if env == "shkadov":
        kw = dict(L0=550.0, n_jets=1)
        kw.update(kwargs)
        e = BatchedEnv("shkadov", batch=1, device=device, seed=seed, init="rest", **kw)
        d = e.cfg.d
        n = d["n_warmup"] if n_warmup is None else int(n_warmup)
        e.reset(n_warm=torch.tensor([n], dtype=torch.int32))                       # all warm actions in one launch
        return {"x": d["x"].copy(), "h": e.get_state("h")[0].cpu().numpy(), "q": e.get_state("q")[0].cpu().numpy()}
    if env == "rayleigh":
        kw = dict(n_sgts=1)
        kw.update(kwargs)
        e = BatchedEnv("rayleigh", batch=1, device=device, seed=seed, init=False, **kw)
        d = e.cfg.d
        n = d["n_warmup"] if n_warmup is None else int(n_warmup)
        e.reset()
        left = n
        while left > 0:
            k = min(chunk, left)
            e.step_fused(torch.zeros(k, 1, e.act_dim, dtype=e.dtype, device=e.device))
            left -= k
        shape = (d["nx"] + 2, d["ny"] + 2)
        return {k: e.get_state(k)[0].cpu().numpy().reshape(shape) for k in ("u", "v", "p", "T")}
    if env == "sloshing":
        e = BatchedEnv("sloshing", batch=1, device=device, seed=seed, init="rest", **kwargs)
        d = e.cfg.d
        n = d["n_warmup"] if n_warmup is None else int(n_warmup)
        e.reset()
        t = np.arange(n) * d["dt_act"]
        acts = torch.as_tensor(sloshing_signal(t).reshape(n, 1, 1), device=e.device)
        for i in range(0, n, chunk):
            e.step_fused(acts[i:i + chunk])
        return {"x": d["x"].copy(), "h": e.get_state("h")[0].cpu().numpy(), "q": e.get_state("q")[0].cpu().numpy()}
    raise ValueError(f"{env}: the reference has no init.py for this env")
