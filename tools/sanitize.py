"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py [envs...]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from beacon_b200 import BatchedEnv

envs = sys.argv[1:] or ["shkadov", "shkadov41", "shkadov20", "shkadov3", "burgers", "sloshing", "lorenz", "vortex", "rayleigh", "mixing"]
rng = np.random.default_rng(0)
for name in envs:
    kw, base = {}, name
    if name == "shkadov41":
        base, kw = "shkadov", dict(n_jets=41, per_jet_rwd=True)
    if name == "shkadov20":
        base, kw = "shkadov", dict(n_jets=20)
    if name == "shkadov3":
        base, kw = "shkadov", dict(n_jets=3)
    e = BatchedEnv(base, batch=3, **kw)
    if base == "shkadov":
        e.reset(n_warm=torch.tensor([0, 1, 2], dtype=torch.int32))
    else:
        e.reset()
    if e.act_is_int:
        a = torch.tensor([0, 1, 2], dtype=torch.int32, device="cuda")
    else:
        a = torch.as_tensor(rng.uniform(-1, 1, (3, e.act_dim)), device="cuda")
    k = 1 if base in ("rayleigh", "mixing") else 2
    for _ in range(k):
        out = e.step(a)
    torch.cuda.synchronize()
    print(name, "ok", float(out[1].sum()), flush=True)
    e.close()
