#!/usr/bin/env python
"""bench.py — env-actions/sec of the batched env dynamics on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--env shkadov] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host cores (oracle port)

A "step" is one gym step of the whole batch (= batch env-actions) through the C-ABI with all
solver sub-steps fused in one kernel launch.  Default workload = BASELINE.json configs[1]:
shkadov-v0, 10 jets, 1024 envs per GPU (weak scaling: every rank steps its own 1024 envs, no
data-path collective).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (ctor kwargs, default batch per GPU, description)
    "shkadov": (dict(n_jets=10), 1024, "shkadov-v0 n_jets=10 nx=1350, 50 sub-steps/action"),
    "shkadov_separable": (dict(n_jets=41, per_jet_rwd=True), 512, "shkadov_separable-v0 n_jets=41 nx=2900"),
    "rayleigh": (dict(), 4096, "rayleigh-v0 50x50, 200 sub-steps/action, Jacobi Poisson"),
    "mixing": (dict(), 1024, "mixing-v0 100x100, 250 sub-steps/action, Jacobi Poisson"),
    "burgers": (dict(), 1, "burgers-v0 nx=500, 62 sub-steps/action, single env"),
    "sloshing": (dict(), 4096, "sloshing-v0 nx=200, 50 sub-steps/action"),
    "lorenz": (dict(), 65536, "lorenz-v0 LSRK4"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--env", default="shkadov", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="envs per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# algorithmic bytes per env-action (SURVEY.md §8d, S-model: every live field read + written once
# per solver sub-step / Jacobi sweep), fp64
# --------------------------------------------------------------------------------------------
def algorithmic_bytes(env_name, cfg, sweeps_per_action=0.0):
    d = cfg.d
    if env_name.startswith("shkadov"):
        return 3200.0 * d["nx"]
    if env_name == "burgers":
        return 3 * d["nx"] * 8.0 * d["ndt_act"]
    if env_name == "sloshing":
        return 8 * (d["nx"] + 2) * 8.0 * d["ndt_act"]
    if env_name == "rayleigh":
        return (21 * d["ndt_act"] + 3 * sweeps_per_action) * (d["nx"] + 2) * (d["ny"] + 2) * 8.0
    if env_name == "mixing":
        return (20 * d["ndt_act"] + 3 * sweeps_per_action) * (d["nx"] + 2) * (d["ny"] + 2) * 8.0
    if env_name == "lorenz":
        return 112.0
    raise ValueError(env_name)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement of the reference, all host threads)
# --------------------------------------------------------------------------------------------
class CpuArm:
    """Steps `B` oracle envs per call with one env per host thread (oracle/beacon_oracle.c)."""

    def __init__(self, env_name, kwargs, B, seed=0):
        import ctypes as C
        from oracle import beacon_oracle as bo
        self.C, self.bo, self.B, self.name = C, bo, B, env_name
        self.lib = bo.lib()
        self.cores = int(self.lib.orc_num_threads())
        self.rng = np.random.default_rng(seed)
        kw = {k: v for k, v in kwargs.items() if k != "per_jet_rwd"}
        base = "shkadov" if env_name.startswith("shkadov") else env_name
        self.proto = bo.ENVS[base](**kw)
        e = self.proto
        e.reset()
        rep = lambda a: np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape)).copy()
        if base == "shkadov":
            self.st = [rep(e.h), rep(e.q), rep(e.rhsh), rep(e.rhsq), np.zeros((B, e.n_jets)), np.zeros((B, e.n_jets))]
            self.obs, self.rwd, self.blow = np.zeros((B, e.n_jets * e.n_obs)), np.zeros(B), np.zeros(B, dtype=np.uint8)
        elif base == "burgers":
            self.st = [rep(e.u), rep(e.up), rep(e.upp)]
            self.obs, self.rwd = np.zeros((B, 5)), np.zeros(B)
        elif base == "sloshing":
            self.st = [rep(e.h), rep(e.q), rep(e.rhsh), rep(e.rhsq), np.zeros(B), np.zeros(B)]
            self.obs, self.rwd = np.zeros((B, e.n_obs)), np.zeros(B)
        elif base == "lorenz":
            self.st = [rep(e.x), rep(e.fx)]
            self.obs, self.rwd = np.zeros((B, 6)), np.zeros(B)
        else:
            scal = e.T if base == "rayleigh" else e.C
            self.st = [rep(e.u), rep(e.v), rep(e.p), rep(scal)]
            self.iters = np.zeros(B, dtype=np.int64)

    def step(self):
        C, e, B, L, rng = self.C, self.proto, self.B, self.lib, self.rng
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        base = "shkadov" if self.name.startswith("shkadov") else self.name
        if base == "shkadov":
            acts, noise = rng.uniform(-1, 1, (B, e.n_jets)), rng.uniform(-e.sigma, e.sigma, (B, e.ndt_act))
            L.orc_shkadov_step_batch(C.byref(e.cfg), B, *[P(a) for a in self.st], P(acts), P(noise), P(self.obs), P(self.rwd), P(self.blow))
        elif base == "burgers":
            acts, noise = rng.uniform(-1, 1, B), rng.uniform(-e.sigma, e.sigma, B)
            L.orc_burgers_step_batch(B, e.nx, C.c_double(e.dx), C.c_double(e.dt), e.ndt_act, e.ctrl_pos, C.c_double(e.amp),
                                     C.c_double(e.u_target), 5, *[P(a) for a in self.st], P(acts), P(noise), P(self.obs), P(self.rwd))
        elif base == "sloshing":
            acts = rng.uniform(-1, 1, B)
            L.orc_sloshing_step_batch(B, e.nx, C.c_double(e.dx), C.c_double(e.dt), e.ndt_act, e.n_interp, C.c_double(e.g),
                                      C.c_double(e.amp), C.c_double(e.alpha), *[P(a) for a in self.st], P(acts), P(self.obs), P(self.rwd))
        elif base == "lorenz":
            acts = rng.integers(0, 3, B).astype(np.int32)
            L.orc_lorenz_step_batch(B, C.c_double(e.sigma), C.c_double(e.rho), C.c_double(e.beta), C.c_double(e.dt), e.ndt_act,
                                    *[P(a) for a in self.st], P(acts), P(self.obs), P(self.rwd))
        elif base == "rayleigh":
            seg = np.stack([e.Th + e.condition(rng.uniform(-1, 1, e.n_sgts)) for _ in range(B)])
            L.orc_mac_solve_batch(C.byref(e.cfg), B, *[P(a) for a in self.st], C.c_double(e.Tc), P(seg), e.n_sgts, e.nx_sgts, None, P(self.iters))
        else:
            wall = np.array([e.get_control(int(a)) for a in rng.integers(0, 4, B)], dtype=np.float64)
            L.orc_mac_solve_batch(C.byref(e.cfg), B, *[P(a) for a in self.st], C.c_double(0.0), None, 0, 0, P(wall), P(self.iters))


def cpu_sample(env_name, kwargs, seconds):
    """Times the oracle on a bounded sample: one env per host thread x a few actions."""
    arm = CpuArm(env_name, kwargs, B=1)
    cores = arm.cores
    per_env = {"shkadov": 8, "shkadov_separable": 4, "rayleigh": 1, "mixing": 1, "burgers": 64, "sloshing": 64, "lorenz": 65536}[env_name]
    B = 1 if env_name == "burgers" else cores * per_env
    arm = CpuArm(env_name, kwargs, B=B)
    arm.step()                                   # warm-up (page faults, thread pool)
    n, t0 = 0, time.perf_counter()
    while True:
        arm.step()
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or (env_name in ("mixing",) and n >= 2):
            break
    return {"value": B * n / el, "unit": "env-actions/s", "cores": cores if B > 1 else 1, "kind": "port",
            "sample": f"{B} envs x {n} actions of the same workload, oracle C port (oracle/beacon_oracle.c), {el:.1f} s"}


# --------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    kwargs, defB, desc = WORKLOADS[args.env]
    steps = args.steps or 5
    arm = CpuArm(args.env, kwargs, B=1)
    cores = arm.cores
    per_env = {"shkadov": 4, "shkadov_separable": 2, "rayleigh": 1, "mixing": 1, "burgers": 1, "sloshing": 32, "lorenz": 8192}[args.env]
    B = 1 if args.env == "burgers" else cores * per_env
    arm = CpuArm(args.env, kwargs, B=B)
    for _ in range(max(1, min(args.warmup, 3))):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.step()
    el = time.perf_counter() - t0
    val = B * steps / el
    line = {"impl": "reference", "metric": "env-actions/sec", "value": val, "unit": "env-actions/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "env": args.env, "sample_envs_per_step": B,
                       "note": "reference algorithm on host cores: oracle C port, one env per thread; each step is a bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": "env-actions/s", "cores": cores if B > 1 else 1, "kind": "port",
                             "sample": f"{B} envs x {steps} actions"},
            "e2e": {"value": val, "unit": "env-actions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from beacon_b200 import BatchedEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kwargs, defB, desc = WORKLOADS[args.env]
    B = args.batch or defB
    K = args.steps or {"shkadov": 400, "shkadov_separable": 100, "rayleigh": 10, "mixing": 3, "burgers": 200, "sloshing": 200, "lorenz": 500}[args.env]
    W = max(args.warmup, 3)
    base = "shkadov" if args.env.startswith("shkadov") else args.env
    env = BatchedEnv(base, batch=B, device=local, seed=1234, env_index_base=rank * B, **kwargs)

    # synthetic inputs, resident in HBM before the timed region
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    n_tot = K + W
    if env.act_is_int:
        hi = 4 if base == "mixing" else 3
        actions = torch.randint(0, hi, (n_tot, B), generator=g, device=dev, dtype=torch.int32)
    else:
        actions = torch.rand(n_tot, B, env.act_dim, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    if base == "shkadov":
        nw = torch.randint(0, 401, (B,), generator=g, device=dev, dtype=torch.int32)
        t0 = time.perf_counter()
        env.reset(n_warm=nw)
        torch.cuda.synchronize()
        reset_s = time.perf_counter() - t0
    else:
        env.reset()
        reset_s = None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # 256 MB > 126 MB L2
    want_iters = base in ("rayleigh", "mixing")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for k in range(W):
        env.step(actions[k], want_iters=want_iters)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sweeps = 0
    launches0 = env.launches
    with ClockSampler(local) as clk:
        barrier()
        for k in range(K):
            flush.zero_()                                    # L2 flush between timed iterations (untimed)
            ev[k][0].record()
            env.step(actions[W + k], want_iters=want_iters)
            ev[k][1].record()
            if want_iters:
                sweeps += int(env.last_iters.sum().item())
        barrier()
    launches = env.launches - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * B * K / (total_ms_max * 1e-3)
    status_bad = int((env.status != 0).sum().item())

    # ---- end-to-end through the public host-buffer API (H2D + step + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        host_act = actions[:K + W].cpu().pin_memory()
        out = env.alloc_host_outputs()
        for k in range(W):
            env.step_host(host_act[k], out=out)
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            env.step_host(host_act[W + k], out=out)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        rb = 8
        e2e = {"value": world * B * K / float(tt.item()), "unit": "env-actions/s",
               "h2d_bytes_per_step": int(host_act[0].numel() * host_act[0].element_size()),
               "d2h_bytes_per_step": int(B * env.n_obs * rb + B * env.rwd_dim * rb + 2 * B + 4 * B)}

    if rank == 0:
        sweeps_per_action = sweeps / (B * K) if want_iters else 0.0
        abytes = algorithmic_bytes(args.env, env.cfg, sweeps_per_action)
        peak, peak_src = measured_peak()
        avg_launch_s = (total_ms / K) * 1e-3
        achieved = abytes * B / avg_launch_s / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu capture of the
        # same command (profiles/traffic.json, written from profiles/*_ncu_<env>.json); per-env traffic is
        # launch-size independent (every env loads and stores its own state once), so a launch of another
        # batch size is scaled by envs
        traffic, secondary = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            t = json.load(open(tp)).get(args.env)
            if t:
                traffic = float(t["bytes_per_launch"]) * B / float(t["grid"])
                if "fp64_pipe_active_pct" in t:     # the physical limiter of the fused kernels (same ncu capture)
                    secondary = {"bound": "fp64 pipe (ncu sm__pipe_fp64_cycles_active, committed capture)",
                                 "frac": t["fp64_pipe_active_pct"] / 100.0, "issue_slots_frac": t.get("issue_active_pct", 0) / 100.0}
        line = {
            "metric": "env-actions/sec", "value": value, "unit": "env-actions/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "env": args.env, "envs_per_gpu": B, "global_batch": B * world,
                       "actions_per_launch": 1, "parallelism": f"env-sharded x{world}, no data-path collective",
                       "l2": "256 MB flush between timed steps", "noise": "on-device Philox",
                       "status_nonzero_envs": status_bad,
                       **({"jacobi_sweeps_per_action": sweeps_per_action} if want_iters else {}),
                       **({"reset_seconds_random_warm_0_400": reset_s} if reset_s is not None else {})},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "secondary": secondary,
                         "model": "S-model algorithmic bytes (SURVEY.md §8d): %.0f B per env-action x %d envs per launch; "
                                  "sub-steps are fused on chip so DRAM traffic is far below this (F-model)" % (abytes, B)},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(args.env, kwargs, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
