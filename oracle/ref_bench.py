"""CPU-baseline timing of the UNMODIFIED reference (numpy + numba) — test / bench infrastructure only.

BASELINE.md §3: one unmodified reference env per host process, P = os.cpu_count() processes, numba warmed
by one untimed step, random actions as in the reference's run.py scripts (shkadov/run.py:12-33,
rayleigh/run.py:12-33), wall clock over a bounded sample.  Only `bench.py`'s cpu_baseline leg and its
`--impl reference` arm call this; the product never does.  The modules come from /root/reference in the
build container, else from the copy staged by oracle/make_ref.py (oracle/_ref/, travels to the GPU box).
"""
import multiprocessing as mp
import os
import time


def available():
    from oracle import refload
    return refload.available()


def _actions(env_name, env, rng):
    import numpy as np
    if env_name.startswith("shkadov"):
        return lambda: rng.uniform(-1.0, 1.0, env.n_jets)
    if env_name == "rayleigh":
        return lambda: rng.uniform(-1.0, 1.0, env.n_sgts)
    if env_name == "mixing":
        return lambda: int(rng.integers(0, 4))
    if env_name == "lorenz":
        return lambda: int(rng.integers(0, 3))
    if env_name == "vortex":
        return lambda: rng.uniform(-1.0, 1.0, 2)
    return lambda: np.array([rng.uniform(-1.0, 1.0)])


def _worker(env_name, kwargs, seconds, idx, barrier, q):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    import warnings
    warnings.simplefilter("ignore")
    import numpy as np
    from oracle import refload
    try:
        base = "shkadov" if env_name.startswith("shkadov") else env_name
        env, mod = refload.make(base, **kwargs)
        if base == "shkadov":
            env.rand_init = False                  # the warm-up steps of reset() are not part of the metric
        env.reset()
        act = _actions(env_name, env, np.random.default_rng(1000 + idx))
        env.step(act())                            # numba JIT + first-touch, untimed
        barrier.wait(timeout=600)
        n, t0 = 0, time.perf_counter()
        while True:
            out = env.step(act())
            n += 1
            if out[2]:
                env.reset()
            el = time.perf_counter() - t0
            if el >= seconds:
                break
        q.put((idx, n, el))
    except Exception as e:                          # never leave the parent waiting
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((idx, -1, repr(e)))


def time_reference(env_name, kwargs, seconds, procs=None):
    """env-actions/s of P unmodified reference envs stepping concurrently for ~`seconds` each."""
    kwargs = {k: v for k, v in kwargs.items() if k != "per_jet_rwd"}
    P = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(P), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(env_name, kwargs, seconds, i, barrier, q), daemon=True) for i in range(P)]
    [p.start() for p in ps]
    res = [q.get(timeout=900) for _ in ps]
    [p.join(timeout=60) for p in ps]
    bad = [r for r in res if r[1] < 0]
    if bad:
        raise RuntimeError(f"reference worker failed: {bad[0][2]}")
    n = sum(r[1] for r in res)
    el = max(r[2] for r in res)
    return {"value": n / el, "unit": "env-actions/s", "cores": P, "kind": "reference",
            "sample": f"{P} processes x 1 unmodified reference env (numpy + numba, JIT warmed by one untimed step), random actions, "
                      f"{n} env-actions in {el:.1f} s"}
