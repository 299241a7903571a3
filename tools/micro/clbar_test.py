"""Cost of a thread-block-cluster barrier per Jacobi sweep, measured on the rayleigh kernel (round 2).

Build the library with -DMAC_CLUSTER_BARRIER_TEST (python -m beacon_b200.build --tag=clbar -DMAC_CLUSTER_BARRIER_TEST):
the kernel is then launched in clusters of 4 CTAs and the per-sweep __syncthreads of its Poisson loop becomes
barrier.cluster.arrive.release + wait.acquire.  ALL envs must take the same number of sweeps (identical actions),
otherwise the CTAs of a cluster deadlock.  Result on one B200, 4096 identical envs, 5452 sweeps per action:
101.5 ms per step with __syncthreads, 134.7 ms with the cluster barrier = +855 cycles per sweep and CTA (a sweep is
~1360 cycles).  Consequence: a mixing kernel that splits an environment over a cluster and synchronises it once
per sweep would spend more on the barrier than it gains from de-phased CTAs.
"""
import sys, torch, time, os
sys.path.insert(0, ".")
from beacon_b200 import BatchedEnv
B = 4096
env = BatchedEnv("rayleigh", batch=B)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
a1 = (torch.rand(6, 1, 10, generator=g, device="cuda", dtype=torch.float64) * 2 - 1).expand(6, B, 10).contiguous()   # identical envs
for k in range(2): env.step(a1[k], want_iters=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(2, 6): env.step(a1[k], want_iters=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 4
print(f"{B * 1e3 / ms:12.0f} env-actions/s  {ms:8.3f} ms/step  sweeps/action {float(env.last_iters.double().mean()):.0f}")
