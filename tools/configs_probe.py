"""Secondary BASELINE configs on one GPU: C3 (active obstacle, 4096 x N=20) throughput and C5 (latency mode: 64 x N=40,
tightened eps_prim) p50 / p99 per-cycle latency measured end to end through the host-buffer C-ABI call."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mpcc_manipulator_b200 as M
from bench import q_home

def make(B, N, overrides=None):
    mpc = M.BatchMPC(B, N); mpc.load_nn(); mpc.set_params(M.load_default_params(overrides=overrides))
    ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
    mpc.set_tracks(M.load_track_json(None, ee))
    return mpc

out = {}
# ---- C3: active moving obstacle (python/main_w_sim.py:21-46 values), 4096 x N=20 ----
B, N = 4096, 20
rng = np.random.default_rng(1)
mpc = make(B, N, {"model.tol_envcol": 1.0, "model.tol_sing": 0.018, "model.desired_ee_velocity": 0.1})
x = np.tile(np.r_[q_home(), 0., 0.], (B, 1)); x[:, :7] += rng.uniform(-0.05, 0.05, (B, 7)); u = np.zeros((B, 8))
obs = np.c_[np.array([0.48, 0.218, 0.521]) + rng.uniform(-0.05, 0.05, (B, 3)), np.full(B, 5.0)]
ts = []
for c in range(14):
    t0 = time.perf_counter(); r = mpc.run_cycle(x, u, obs, want_horizon=False); dt = time.perf_counter() - t0
    if c >= 4: ts.append(dt)
    u = r["u0"]; x = mpc.sim_time_step(r["x0"], u); obs[:, 2] += 0.05 * 0.01
st = mpc.stats()
out["C3"] = {"batch": B, "horizon": N, "e2e_ms_per_cycle": 1e3 * float(np.mean(ts)), "e2e_solves_per_s": B / float(np.mean(ts)), "last_cycle": st}
mpc.close()
# ---- C5: latency mode ----
B, N = 64, 40
rng = np.random.default_rng(3)
mpc = make(B, N, {"sqp.eps_prim": 0.01})  # 2-4 SQP iterations per cycle (1e-3 never terminates: every cycle runs max_iter)
x = np.tile(np.r_[q_home(), 0., 0.], (B, 1)); x[:, :7] += rng.uniform(-0.05, 0.05, (B, 7)); u = np.zeros((B, 8))
ts, its = [], []
for c in range(220):
    t0 = time.perf_counter(); r = mpc.run_cycle(x, u, want_horizon=False); dt = time.perf_counter() - t0
    if c >= 20: ts.append(dt); its.append(r["iters"].mean())
    u = r["u0"]; x = mpc.sim_time_step(r["x0"], u)
ts = np.array(ts) * 1e3
out["C5"] = {"batch": B, "horizon": N, "eps_prim": 0.01, "cycles": len(ts), "p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)),
             "max_ms": float(ts.max()), "mean_sqp_iters": float(np.mean(its)), "solved_last": mpc.stats()["solved"]}
mpc.close()
print(json.dumps(out))
