"""Kernel-variant sweep for shkadov (tuning aid): env-actions/s per (C, T, MINB) variant.
Usage (GPU box): python tools/sweep_shkadov.py [n_jets] [batch ...]"""
import os
import subprocess
import sys

CODE = r'''
import sys, torch, time
sys.path.insert(0, ".")
from beacon_b200 import BatchedEnv
nj, B = int(sys.argv[1]), int(sys.argv[2])
env = BatchedEnv("shkadov", batch=B, n_jets=nj, seed=1)
env.reset()
K = 30
acts = torch.rand(K + 3, B, nj, device="cuda", dtype=torch.float64) * 2 - 1
for k in range(3): env.step(acts[k])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K): env.step(acts[3 + k])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"{B * 1e3 / ms:12.0f} env-actions/s  {ms:8.3f} ms/step")
'''
nj = int(sys.argv[1]) if len(sys.argv) > 1 else 10
batches = [int(x) for x in sys.argv[2:]] or [1024, 4096]
cfgs = os.environ["SWEEP_CFGS"].split(";") if os.environ.get("SWEEP_CFGS") else ["6,256,2", "10,192,2", "6,512,1"] if nj <= 12 else ["10,192,2", "6,512,1", "12,512,1"]
for cfg in cfgs:
    for B in batches:
        env = dict(os.environ, BEACON_SHKADOV_CFG=cfg)
        r = subprocess.run([sys.executable, "-c", CODE, str(nj), str(B)], env=env, capture_output=True, text=True)
        print(f"cfg={cfg:10s} n_jets={nj} B={B:5d}: {(r.stdout.strip() or r.stderr.strip()[-200:])}", flush=True)
