#!/usr/bin/env python
"""Benchmark of the batched MPCC control cycle (BASELINE.json metric: batched MPCC SQP solves/sec, 4096 x N=20).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (CPU oracle)

One "step" = one closed-loop control cycle (mpcc::MPC::runMPC for every instance of the batch) followed by the
plant step that produces the next cycle's inputs.  Workload = BASELINE configs[1] ("C2" in SURVEY.md 8d): 4096
Panda instances per GPU, perturbed initial joint angles (seeded), same track, N = 20.  Prints ONE JSON line.

`value`: device-resident inputs, CUDA events on the library's stream around every step, L2 flushed between steps.
`e2e`  : the same cycle through the C-ABI call that takes HOST buffers (mpcc_cuda_run_cycle), pinned memory,
         host->device and device->host copies inside the timed region.
`roofline`: dominant kernel's algorithmic FP64 FLOPs / its live CUDA-event duration, against the FP64 FMA peak
         measured in this run (MEASURED_PEAKS.json has no FP64 figure; the path is fp64 and tcgen05 has no f64 kind).
`cpu_baseline`: the CPU oracle (restated reference algorithm) on a bounded sample, this box's host cores.
"""
import argparse
import json
import os
import select
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "batched MPCC SQP solves/sec (4096xN=20)"
UNIT = "solves/s"
# algorithmic FLOPs: minimal MAC counts of both networks with 7 forward-mode tangent columns
MLP_FLOP_PER_STAGE = 2.0 * 1746688
# what the default kernel executes since the self net runs in reverse mode (value forwards + one adjoint sweep: 43 605 instead of 142 336 MAC per sample);
# the roofline keeps SURVEY's figure (the contract's per-unit work), this one is reported next to it
MLP_EXECUTED_FLOP_PER_STAGE = 2.0 * (1604352 + 43605)
KIN_FLOP_PER_STAGE = 2.0 * 9000
# SQP kernel, per interior-point iteration and stage: Riccati factorisation 4.2 k MAC (polytopic rank-11
# update 1155, L^-1 Mnx 576, cost-to-go update 2048, blocks / Cholesky ~400), two sweeps 2 x 384, two gradients ~500,
# three constraint passes ~500
QP_FLOP_PER_STAGE_ITER = 2.0 * 6000
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12
# int8-split MLP kernel (csrc/mlp_oz_kernel.cuh): per 8-sample tile the three 256 x 256 env layers issue 2 passes x S (S + 1) / 2 digit
# products x 8 k-steps of tcgen05.mma 128 x 64 x 32 (S = 7 digits)
OZ_S = 7
OZ_INT8_MAC_PER_TILE = 3 * 2 * (OZ_S * (OZ_S + 1) // 2) * 8 * (128 * 64 * 32)
MLP_FP64_KERNEL_FLAG = 8   # mpcc_cuda_config.reserved bit 3: the fp64 DMMA kernel instead


def q_home():
    return np.array([0, 0, 0, -np.pi / 2, 0, np.pi / 2, np.pi / 4])


def synthetic_inputs(B, seed):
    """C2: q0 = q_home + U(-0.05, 0.05)^7, s0 = vs0 = 0, u0 = 0."""
    rng = np.random.default_rng(seed)
    x0 = np.tile(np.r_[q_home(), 0.0, 0.0], (B, 1))
    x0[:, :7] += rng.uniform(-0.05, 0.05, (B, 7))
    return x0, np.zeros((B, 8))


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU side: persistent oracle workers (one process per core), driven step by step over pipes
# ------------------------------------------------------------------------------------------------
def cpu_worker_main(args):
    """Child process: owns `per` oracle MPC instances (warm starts persist across steps like on the GPU).
    Protocol on stdin/stdout: 'step' -> one closed-loop cycle for every instance -> 'done <seconds>'."""
    from oracle import oracle as O
    nn = O.OracleNN()
    ee = O.fk(O.Q_HOME)[0]
    X, Y, Z, R = O.load_track()
    X, Y, Z = O.shift_track(X, Y, Z, ee)
    x0, u0 = synthetic_inputs(args.per, 1000 + args.worker_id)
    inst = []
    for b in range(args.per):
        o = O.OracleMPC(N=args.horizon, nn=nn)
        o.set_track(X, Y, Z, R)
        inst.append([o, x0[b], u0[b]])
    print("ready", flush=True)
    for line in sys.stdin:
        if line.strip() != "step":
            break
        t0 = time.perf_counter()
        for it in inst:
            r = it[0].run(it[1], it[2])
            it[2] = r["u0"]
            it[1] = O.sim_time_step(r["x0"], r["u0"], 0.01)
        print(f"done {time.perf_counter() - t0:.6f}", flush=True)


class CpuPool:
    def __init__(self, workers, per, horizon):
        self.workers, self.per = workers, per
        env = dict(os.environ, OMP_NUM_THREADS="1")
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
            env.pop(k, None)
        self.procs = [subprocess.Popen([sys.executable, str(ROOT / "bench.py"), "--cpu-worker", "--worker-id", str(i), "--per", str(per), "--horizon", str(horizon)],
                                       stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env) for i in range(workers)]
        for p in self.procs:
            assert p.stdout.readline().strip() == "ready"

    def step(self):
        """One cycle for every instance of every worker; returns wall seconds."""
        t0 = time.perf_counter()
        for p in self.procs:
            p.stdin.write("step\n"); p.stdin.flush()
        for p in self.procs:
            assert p.stdout.readline().startswith("done")
        return time.perf_counter() - t0

    def close(self):
        for p in self.procs:
            try:
                p.stdin.close(); p.wait(timeout=5)
            except Exception:
                p.kill()


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle")])


def cpu_baseline(horizon, budget_s=14.0):
    """Bounded sample on the host cores: one process per core, each owning 2 instances; cold cycle untimed."""
    build_oracle()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # single thread first (alone on the machine)
    one = CpuPool(1, 2, horizon)
    one.step()
    t_single, n_single = 0.0, 0
    while t_single < budget_s * 0.3:
        t_single += one.step(); n_single += 2
    one.close()
    pool = CpuPool(cores, 2, horizon)
    pool.step()
    t, n = 0.0, 0
    while t < budget_s * 0.7:
        t += pool.step(); n += cores * 2
    pool.close()
    return {"value": n / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} warm-started closed-loop solves (N={horizon}) of the C2 workload over {cores} processes x 2 instances, oracle/liborc.so (-O2), first (cold) cycle untimed",
            "single_thread": {"value": n_single / t_single, "cores": 1, "sample": f"{n_single} solves"}}


def run_reference(args):
    """--impl reference: the reference algorithm (CPU oracle port; the reference itself cannot be built here)
    on every host core.  A step = one closed-loop cycle of a bounded sample batch (cores x 2 instances)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    build_oracle()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per = 2
    pool = CpuPool(cores, per, args.horizon)
    for _ in range(max(args.warmup, 1)):
        pool.step()
    ts = [pool.step() for _ in range(args.steps)]
    pool.close()
    total = float(np.sum(ts))
    n = cores * per * args.steps
    val = n / total
    sample = f"{cores * per} instances per step ({cores} processes x {per}), N={args.horizon}, warm-started closed loop"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "C2: 4096 Panda instances, perturbed q0, default track, N=20 (bounded sample per step)", "horizon": args.horizon, "sample": sample},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class DevPtr:
    """Wrap a raw device pointer of the library as a __cuda_array_interface__ object (torch.as_tensor reads it)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


# params record offsets (include/mpcc_cuda.h): model7 | cost12 | ...
P_COST = 7
P_QC, P_QL, P_QVS, P_QORI = P_COST + 0, P_COST + 2, P_COST + 3, P_COST + 4

CONFIGS = {
    # name: (batch per GPU, horizon, description)
    "c2": (4096, 20, "C2 (BASELINE configs[1]): 4096 Panda instances per GPU, q0 = q_home + U(-0.05,0.05) seeded, default track.json, N=20, closed loop"),
    "c3": (4096, 20, "C3 (BASELINE configs[2]): 4096 instances per GPU, N=20, live moving obstacle (env-collision rows active), tol_envcol=1, tol_sing=0.018, v_des=0.1; "
                     "fixed-base Panda (the reference has no mobile-base model, SURVEY F4)"),
    "c4": (8192, 20, "C4 (BASELINE configs[3]): heterogeneous batch, 8192 instances per GPU (65536 on 8), per-instance track.py-family spline (fitted on the device) "
                     "and per-instance cost weights qC/qL/qOri/qVs, N=20"),
    "c5": (64, 40, "C5 (BASELINE configs[4]): latency mode, 64 instances, N=40, eps_prim=0.01 (several SQP iterations per cycle), per-cycle latency through the host-buffer call"),
}


def setup_config(M, name, B, N, local, rank):
    """Build the handle and the synthetic closed-loop start of one BASELINE configuration."""
    over = None
    if name == "c3":
        over = {"model.tol_envcol": 1.0, "model.tol_sing": 0.018, "model.desired_ee_velocity": 0.1}
    if name == "c5":
        over = {"sqp.eps_prim": 0.01}
    mpc = M.BatchMPC(B, N, device=local, qp_eps=float(os.environ.get("MPCC_BENCH_QP_EPS", "0")), flags=int(os.environ.get("MPCC_BENCH_FLAGS", "0")))  # 0 = library default; env: diagnostic sweeps only
    mpc.load_nn()
    base = M.load_default_params(overrides=over)
    seed = {"c2": 0, "c3": 1, "c4": 2, "c5": 3}[name] + 1000 * rank
    rng = np.random.default_rng(seed)
    x0 = np.tile(np.r_[q_home(), 0.0, 0.0], (B, 1))
    x0[:, :7] += rng.uniform(-0.05, 0.05, (B, 7))
    u0 = np.zeros((B, 8))
    obs = None
    info = {}
    if name == "c4":
        P = np.tile(base, (B, 1))
        P[:, P_QC] = rng.uniform(200, 1000, B); P[:, P_QL] = rng.uniform(50, 200, B); P[:, P_QORI] = rng.uniform(10, 100, B); P[:, P_QVS] = rng.uniform(5, 40, B)
        mpc.set_params(P)
    else:
        mpc.set_params(base)
    # tracks start at the EE position at q_home, as main.cpp does (track.cpp:58-60); FK through the library itself
    ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
    if name == "c4":
        t = np.linspace(np.pi / 2, 5 * np.pi / 2, 100)
        a, b, c = rng.uniform(1.5, 3, B), rng.uniform(1.5, 3, B), rng.uniform(0, 2.5, B)
        X = 0.1 * a[:, None] * np.sin(t)[None]; Y = 0.1 * b[:, None] * np.sin(2 * t)[None]; Z = 0.1 * c[:, None] * np.cos(t)[None]
        X = X - X[:, :1] + ee[0]; Y = Y - Y[:, :1] + ee[1]; Z = Z - Z[:, :1] + ee[2]
        R = np.tile(np.diag([1.0, -1.0, -1.0]).ravel(), (B, 100, 1))
        t0 = time.perf_counter()
        mpc.fit_tracks_device(X, Y, Z, R, np.arange(B))     # ArcLengthSpline::fitSpline for every instance, on the device
        info["track_fit_ms"] = 1e3 * (time.perf_counter() - t0)
        x0[:, :7] = q_home()[None] + rng.uniform(-0.03, 0.03, (B, 7))
    else:
        mpc.set_tracks(M.load_track_json(None, ee))
    if name == "c3":
        obs = np.c_[np.array([0.48, 0.218, 0.521]) + rng.uniform(-0.05, 0.05, (B, 3)), np.full(B, 5.0)]
    return mpc, x0, u0, obs, info


def measure(M, name, args, steps, warmup, world, rank, local, dist, full):
    """Closed-loop measurement of one configuration on this rank's GPU.
    Returns the device-resident timing (`value`), the end-to-end timing through the host-buffer call (`e2e`) and, with `full`,
    per-kernel times / stats for the roofline."""
    import torch
    import ctypes as C
    from mpcc_manipulator_b200 import capi
    B, N, desc = CONFIGS[name]
    if name == args.config:
        B = args.batch or B; N = args.horizon or N
    S = N + 1
    Ts = 0.01
    mpc, x_host, u_host, obs_host, info = setup_config(M, name, B, N, local, rank)
    dev = f"cuda:{local}"
    if world > 1:
        # the library's own communicator (NCCL bound at run time): rank 0 creates the id, torch.distributed only carries it
        uid = torch.from_numpy(M.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        mpc.comm_init(uid.cpu().numpy(), rank, world)
    stream = torch.cuda.ExternalStream(mpc.stream, device=local)
    p_u, p_hor, p_st, p_it, p_ok = mpc.result_pointers()
    with torch.cuda.stream(stream):
        x = torch.from_numpy(x_host).to(dev)
        xn = torch.empty_like(x)
        u = torch.from_numpy(u_host).to(dev)
        obs = torch.from_numpy(obs_host).to(dev) if obs_host is not None else None
        u_out = torch.as_tensor(DevPtr(p_u, (B, 8), "<f8"), device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream.synchronize()
    launches = {"n": 0}

    def step():
        """device-resident closed-loop step: cycle -> (enqueue the gather on the side stream) -> plant step"""
        nonlocal x, xn
        mpc.run_cycle_device(x.data_ptr(), u.data_ptr(), obs.data_ptr() if obs is not None else None)
        if world > 1:
            mpc.gather_results()   # the path's only collective: u0 / status / iterations of every rank (NCCL all-gather, off the critical path)
        launches["n"] += mpc.launch_count()
        with torch.cuda.stream(stream):
            u.copy_(u_out)
            if obs is not None:
                obs[:, 2] += 0.05 * Ts     # obstacle moving at 0.05 m/s in z (python/main_w_sim.py:126-129)
        mpc.sim_time_step_device(x.data_ptr(), u.data_ptr(), xn.data_ptr())
        launches["n"] += 1
        x, xn = xn, x

    # ---- settle: the closed loop is run from its cold start until every instance tracks with one SQP iteration per cycle
    #      ("report steady-state cycles/s x B; also a cold first cycle").  The start-up transient is timed on the
    #      way and reported separately (`cold_start`), including round 1's window (cycles 5..24 after the cold start). ----
    settle = args.settle if name != "c5" else 0
    cold = None
    if settle > 0:
        sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(settle)]
        for i in range(settle):
            with torch.cuda.stream(stream):
                sev[i][0].record(stream)
            step()
            with torch.cuda.stream(stream):
                sev[i][1].record(stream)
        mpc.synchronize()
        sms = np.array([a.elapsed_time(b) for a, b in sev])
        cold = {"first_cycle_ms": float(sms[0]), "settle_cycles": settle}
        if settle >= 25:
            cold["cycles_5_24_mean_ms"] = float(sms[5:25].mean())
            cold["value_cycles_5_24"] = world * B / (float(sms[5:25].mean()) * 1e-3)
            cold["note"] = "per-rank device time; cycles_5_24 is the window round 1's bench line was measured on"
    for _ in range(warmup):
        step()
    mpc.synchronize()
    mpc.set_profiling(True)
    sampler = ClockSampler(local) if (full and rank == 0) else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ktimes = np.zeros((steps, 4))
    stats_acc = {}
    qp_iters_acc = 0
    launches["n"] = 0
    t_wall0 = time.perf_counter()
    for i in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            ev[i][0].record(stream)
        step()
        with torch.cuda.stream(stream):
            ev[i][1].record(stream)
        ktimes[i] = mpc.kernel_times()
        if full or i == steps - 1:
            stats_acc = mpc.stats()
            qp_iters_acc += stats_acc["qp_iters"]
    mpc.synchronize()
    if world > 1:
        gathered = mpc.read_gathered()   # waits for the last enqueued gather: the timed region ends with every rank's results everywhere
        assert gathered["u0"].shape[0] == world * B
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    total_ms = float(step_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms_max = float(t.item())
    else:
        total_ms_max = total_ms
    value = world * B * steps / (total_ms_max * 1e-3)
    mpc.set_profiling(False)

    # ---- e2e: HOST buffers through mpcc_cuda_run_cycle (pinned), copies inside the timed region, MPCReturn complete
    #      (u0, the whole mpc_horizon, status, iterations, ok) and, on several GPUs, the gather read back on the host ----
    hx = torch.empty((B, 9), dtype=torch.float64).pin_memory(); hu = torch.empty((B, 8), dtype=torch.float64).pin_memory()
    huo = torch.empty((B, 8), dtype=torch.float64).pin_memory(); hhor = torch.empty((B, S, 17), dtype=torch.float64).pin_memory()
    hobs = torch.from_numpy(obs_host.copy()).pin_memory() if obs_host is not None else None
    hst = torch.empty(B, dtype=torch.int32).pin_memory(); hit = torch.empty(B, dtype=torch.int32).pin_memory(); hok = torch.empty(B, dtype=torch.int32).pin_memory()
    # the same closed loop goes on from where the device-resident measurement left it (warm starts stay on the device)
    hx.copy_(x.cpu()); hu.copy_(u.cpu())
    if hobs is not None:
        hobs.copy_(obs.cpu())
    gu = np.zeros((world * B, 8)); gs = np.zeros(world * B, np.int32); gi = np.zeros(world * B, np.int32)

    def e2e_step():
        rc = capi.lib().mpcc_cuda_run_cycle(mpc.h, C.c_void_p(hx.data_ptr()), C.c_void_p(hu.data_ptr()), C.c_void_p(hobs.data_ptr()) if hobs is not None else None,
                                            C.c_void_p(huo.data_ptr()), C.c_void_p(hhor.data_ptr()), C.c_void_p(hst.data_ptr()), C.c_void_p(hit.data_ptr()), C.c_void_p(hok.data_ptr()))
        if rc != 0:
            raise RuntimeError(capi.lib().mpcc_cuda_last_error().decode())
        if world > 1:
            mpc.gather_results()
            capi._check(capi.lib().mpcc_cuda_read_gathered(mpc.h, capi._p(gu), capi._p(gs), capi._p(gi)))
        # host-side plant step (exact for the linear model; integrator.cpp:55-68), next cycle's inputs
        xs, us = hx.numpy(), huo.numpy()
        xs[:, :7] += Ts * us[:, :7]
        xs[:, 7] += Ts * xs[:, 8] + 0.5 * Ts * Ts * us[:, 7]
        xs[:, 8] += Ts * us[:, 7]
        hu.copy_(huo)
        if hobs is not None:
            hobs[:, 2] += 0.05 * Ts

    e2e_steps = steps if name == "c5" else max(3, min(steps, 20))
    for _ in range(warmup):
        e2e_step()
    if world > 1:
        dist.barrier()
    lat = np.zeros(e2e_steps)
    iters_max = np.zeros(e2e_steps, int)
    hit_np = hit.numpy()   # (a first torch reduction on the pinned tensor costs milliseconds of one-time set-up: keep the bookkeeping in numpy)
    hit_np.max()
    inst_ms = []
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ta = time.perf_counter()
        e2e_step()
        lat[i] = (time.perf_counter() - ta) * 1e3
        iters_max[i] = int(hit_np.max())
        if name == "c5":   # per-instance solve times of this cycle (device timer around each instance's SQP loop), read outside the timed call
            inst_ms.append(mpc.compute_time()[:, 0] * 1e3)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_val = world * B * e2e_steps / t_e2e
    h2d = B * (9 + 8) * 8 + (B * 4 * 8 if hobs is not None else 0)
    d2h = B * 9 * 8 + B * 8 * 8 + B * S * 17 * 8 + 3 * B * 4 + (world * B * 72 if world > 1 else 0)
    res = {"name": name, "workload": desc, "B": B, "N": N, "value": value, "ms_per_step": total_ms_max / steps, "step_ms": step_ms, "ktimes": ktimes,
           "stats": stats_acc, "qp_iters_mean": qp_iters_acc / max(1, steps if full else 1), "launches": launches["n"], "clocks": clocks, "t_wall": t_wall,
           "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                   "api": "mpcc_cuda_run_cycle (host buffers, pinned; u0 + full mpc_horizon + status/iters/ok copied back"
                          + ("; + mpcc_cuda_gather_results / read_gathered over NCCL" if world > 1 else "") + ")",
                   "latency_ms": {"p50": float(np.percentile(lat, 50)), "p90": float(np.percentile(lat, 90)), "p99": float(np.percentile(lat, 99)), "max": float(lat.max())},
                   "max_sqp_iters_per_cycle": {"p50": int(np.percentile(iters_max, 50)), "p99": int(np.percentile(iters_max, 99)), "max": int(iters_max.max())}},
           "info": info, "cold_start": cold}
    if inst_ms:
        a = np.concatenate(inst_ms)
        res["e2e"]["per_instance_sqp_ms"] = {"p50": float(np.percentile(a, 50)), "p90": float(np.percentile(a, 90)), "p99": float(np.percentile(a, 99)), "max": float(a.max()),
                                             "what": "SQP time of each single instance (ComputeTime::total analogue) over all instances and cycles of the e2e run: the batch call returns with its "
                                                     "slowest instance, and which instances wander into the reference algorithm's 30..100-iteration regime at eps_prim = 0.01 is chaotic "
                                                     "(it changes with the last bit of RobotData), so the per-cycle percentiles above move from build to build while these do not"}
    mpc.close()
    return res


def summary_of(r):
    """compact record of a secondary configuration for the JSON line"""
    sm = r["step_ms"]
    return {"workload": r["workload"], "batch_per_gpu": r["B"], "horizon": r["N"], "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
            "device_step_ms": {"p50": float(np.percentile(sm, 50)), "p99": float(np.percentile(sm, 99)), "max": float(sm.max())},
            "e2e": r["e2e"], "kernels_ms": {n: round(float(t), 4) for n, t in zip(KERNEL_NAMES, r["ktimes"].mean(axis=0))},
            "last_step_stats": r["stats"], "cold_start": r["cold_start"], **r["info"]}


KERNEL_NAMES = ["k_prologue", "k_kin", "k_mlp", "k_sqp_warp"]   # the third slot is the MLP launch: k_mlp_oz by default, k_mlp with MPCC_BENCH_FLAGS=8


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mpcc_manipulator_b200 as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback for the product path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = measure(M, args.config, args, args.steps, args.warmup, world, rank, local, dist, full=True)
    B, N, S = r["B"], r["N"], r["N"] + 1
    peak = M.fp64_peak(local) if rank == 0 else None   # before the secondary runs: nothing the metric's line needs depends on them
    # secondary BASELINE configurations, short runs, reported inside the same JSON line (the driver only runs the default command)
    secondary = {}
    if args.config == "c2" and not args.no_secondary:
        names = ["c3", "c4", "c5"] if world == 1 else ["c4"]
        for nm in names:
            st, wu = (300, 20) if nm == "c5" else (10, 3)
            try:   # a failure in a short secondary run must not take the metric's line with it (single GPU only: ranks stay in step)
                secondary[nm] = summary_of(measure(M, nm, args, st, wu, world, rank, local, dist, full=False))
            except Exception as ex:
                if world > 1:
                    raise
                secondary[nm] = {"error": repr(ex)[:300]}

    out = None
    if rank == 0:
        step_ms, ktimes, stats_acc = r["step_ms"], r["ktimes"], r["stats"]
        km = ktimes.mean(axis=0)
        names = KERNEL_NAMES
        dom = int(np.argmax(km))
        peak_src = ("FP64 microbenchmark measured in this run (mpcc_cuda_fp64_peak: larger of the DFMA and the mma.m8n8k4.f64 figure, one shared pipe); MEASURED_PEAKS.json holds no FP64 figure; "
                    "nominal 148 SM x 64 FMA/clk x 1.965 GHz = %.1f" % NOMINAL_FP64_TFLOPS)
        # DRAM traffic per launch: ncu --set full captures of STEADY-STATE launches, committed with their capture conditions;
        # reported as captured (bytes per launch), never divided by this run's kernel time
        traffic = {}
        try:
            traffic = json.loads((ROOT / "profiles" / "dram_traffic.json").read_text())
        except Exception:
            pass
        mlp_flop = B * S * MLP_FLOP_PER_STAGE
        sqp_flop = r["qp_iters_mean"] * N * QP_FLOP_PER_STAGE_ITER
        kernels = {n: round(float(t), 4) for n, t in zip(names, km)}
        share = {n: round(float(t / step_ms.mean()), 4) for n, t in zip(names, km)}

        def roof_of(kernel, flop, ms, note, bound):
            ach = flop / (ms * 1e-3) / 1e12
            tr = traffic.get(kernel)
            return {"bound": bound, "pipe": "fp64", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": tr.get("dram_bytes_per_launch") if isinstance(tr, dict) else tr, "traffic_capture": tr if isinstance(tr, dict) else None,
                    "algorithmic_flops_per_launch": flop, "kernel_ms": float(ms), "peak_source": peak_src, "note": note}
        mlp_fp64_kernel = (int(os.environ.get("MPCC_BENCH_FLAGS", "0")) & MLP_FP64_KERNEL_FLAG) != 0
        if mlp_fp64_kernel:
            roof_mlp = roof_of("k_mlp", mlp_flop, km[2], "dense fp64 contraction (both networks + 7 forward-mode tangents) on mma.sync.m8n8k4.f64 (the fp64 kernel, config.reserved bit 3); "
                               "DMMA and DFMA share one FP64 pipe, which is the roof", "tensor")
        else:
            # default kernel: the three 256 x 256 env layers (88 % of the MACs) as exact int8 digit products on tcgen05 (kind::i8, TMEM accumulators);
            # the same algorithmic fp64 FLOPs are reported against the FP64 pipe (the roof of the fp64 formulation: above 1.0 means the contraction
            # has left that pipe), and the int8 work actually issued against the tensor peak
            roof_mlp = roof_of("k_mlp_oz", mlp_flop, km[2], "both networks with their joint Jacobians (env net: 7 forward-mode tangents; self net: reverse mode); the three 256 x 256 env layers run as exact int8 digit products (7 digits of 7 bits per operand, "
                               "28 products per layer) on tcgen05.mma kind::i8 with TMEM accumulators and recombine in int64 -- fp64-equivalent results (1e-13 of the fp64 kernel). `achieved` / `frac` "
                               "are the ALGORITHMIC fp64 FLOPs against the measured FP64 pipe peak: the roof of the fp64 formulation, exceeded because the contraction no longer runs there; "
                               "`tensor_int8` is the int8 work actually issued against the tensor-core peak", "tensor")
            roof_mlp["executed_flops_per_launch"] = B * S * MLP_EXECUTED_FLOP_PER_STAGE
            roof_mlp["executed_note"] = ("`algorithmic_flops_per_launch` is SURVEY 8d's minimal forward-mode count (1 746 688 MAC per sample); the kernel executes 1 647 957: "
                                         "its self net runs in reverse mode (one adjoint sweep instead of seven tangent columns)")
            tiles = -(-(B * S) // 8)
            int8_ops = 2.0 * OZ_INT8_MAC_PER_TILE * tiles
            bf16_burst = bf16_sust = None
            try:
                mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
                bf16_burst, bf16_sust = mp.get("bf16_tflops"), mp.get("bf16_tflops_sustained")
            except Exception:
                pass
            ach = int8_ops / (km[2] * 1e-3) / 1e12
            pk = 2.0 * bf16_sust if bf16_sust else 4500.0   # int8 dense = 2 x bf16 dense on B200 (4.5 vs 2.25 POP/s nominal); kernel timed inside a long step -> sustained figure
            roof_mlp["tensor_int8"] = {"achieved": ach, "peak": pk, "unit": "TOP/s", "frac": ach / pk, "int8_ops_per_launch": int8_ops,
                                       "peak_source": ("2 x bf16_tflops_sustained of MEASURED_PEAKS.json (of measured; int8 dense is twice bf16 dense on B200)" if bf16_sust
                                                       else "nominal 4.5 POP/s (MEASURED_PEAKS.json absent: of fallback)"),
                                       "frac_of_nominal_4500": ach / 4500.0,
                                       "note": "the MMAs are 128 x 64 x 32 (N = 64 is what eight samples x eight columns give and what 7 accumulators leave room for in TMEM); "
                                               "measured alone such an MMA stream reaches 43 cycles per MMA = 0.74 of the tensor floor (tools/probes/oz_umma_probe.cu); the rest of the kernel "
                                               "time is the fp64 <-> digit conversion, the non-split layers and the phase barriers (DESIGN.md 3)"}
        roof_sqp = roof_of("k_sqp_warp", sqp_flop, km[3], "interior-point / Riccati SQP loop: dependent small factorisations, latency-bound "
                           "(FLOP model x measured interior-point iterations, mean over the timed steps)", "latency (reported against the fp64 pipe)")
        roof = dict(roof_sqp if dom == 3 else roof_mlp)
        roof["dominant_kernel"] = roof_mlp["kernel"] if dom == 2 else names[dom]
        roof["kernel_share_of_step"] = share
        whole = (B * S * (MLP_FLOP_PER_STAGE + KIN_FLOP_PER_STAGE)) / (step_ms.mean() * 1e-3) / 1e12
        roof["whole_cycle"] = {"achieved": whole, "frac": whole / peak, "note": "fixed algorithmic FLOPs of the cycle (networks + kinematics, SURVEY 8d) / mean step time"}
        hbm_peak = None
        try:
            hbm_peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
        except Exception:
            pass
        alg_bytes = B * ((9 + 8 + 4) * 8 + 2 * (17 * N + 9) * 8 + 64 + 16)
        roof["hbm"] = {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / (step_ms.mean() * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                       "note": "arithmetic intensity ~1e4 FLOP/B: HBM fraction is tiny by construction (SURVEY 8d)"}
        out = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": ("f64" if (int(os.environ.get("MPCC_BENCH_FLAGS", "0")) & MLP_FP64_KERNEL_FLAG) else "f64 (three of the MLP layers as exact int8 digit products on tcgen05, recombined in int64: fp64-equivalent, 1e-13 of the fp64 kernel)"), "data": "synthetic",
               "config": {"workload": r["workload"], "name": args.config,
                          "batch_per_gpu": B, "horizon": N, "l2": "flushed (256 MiB memset) between timed steps",
                          "regime": f"steady-state closed-loop tracking: {args.settle} settle cycles from the cold start before the warm-up steps (SURVEY 8d C2); the start-up transient is in `cold_start`",
                          "parallelism": f"dp{world} (instances sharded; all-gather of u0/status/iters only, NCCL on a side stream through mpcc_cuda_gather_results)"},
               "latency_ms": {"p50": float(np.percentile(step_ms, 50)), "p99": float(np.percentile(step_ms, 99)), "max": float(step_ms.max()),
                              "what": "device time of one closed-loop step of the whole batch (CUDA events)"},
               "clocks": r["clocks"], "e2e": r["e2e"],
               "gpu_launches": r["launches"], "kernels_ms": kernels, "roofline": roof, "roofline_mlp": roof_mlp, "roofline_sqp": roof_sqp,
               "last_step_stats": stats_acc, "wall_s_timed_region": r["t_wall"], "cold_start": r["cold_start"]}
        if secondary:
            out["secondary_configs"] = secondary
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(N)
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE configuration (SURVEY 8d); c2 is the metric's")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (0: the configuration's own)")
    ap.add_argument("--horizon", type=int, default=0, help="N (0: the configuration's own)")
    ap.add_argument("--settle", type=int, default=40, help="closed-loop cycles run from the cold start before warm-up (0: time the start-up transient itself)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short C3 / C4 / C5 runs that the default (c2) line carries")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-worker", action="store_true")
    ap.add_argument("--worker-id", type=int, default=0)
    ap.add_argument("--per", type=int, default=2)
    args = ap.parse_args()
    if args.cpu_worker:
        args.horizon = args.horizon or 20
        return cpu_worker_main(args)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        args.horizon = args.horizon or CONFIGS[args.config][1]
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    main()
