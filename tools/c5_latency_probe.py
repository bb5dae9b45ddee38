"""Latency mode (C5: 64 instances, N = 40): where a cycle's time goes -- per-kernel event times, per-instance SQP times, iteration counts."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mpcc_manipulator_b200 as M
from bench import q_home
B, N = 64, 40
rng = np.random.default_rng(3)
mpc = M.BatchMPC(B, N); mpc.load_nn(); mpc.set_params(M.load_default_params(overrides={"sqp.eps_prim": 0.01}))
ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee)); mpc.set_profiling(True)
x = np.tile(np.r_[q_home(), 0., 0.], (B, 1)); x[:, :7] += rng.uniform(-0.05, 0.05, (B, 7)); u = np.zeros((B, 8))
for c in range(60):
    r = mpc.run_cycle(x, u, want_horizon=False)
    kt = mpc.kernel_times(); ct = mpc.compute_time() * 1e3; qi, qf = mpc.qp_counters(); it = r["iters"]
    if c >= 20 and c % 4 == 0:
        w = int(np.argmax(ct[:, 0]))
        print(f"cycle {c}: k_mlp {kt[2]:.2f} k_sqp {kt[3]:.2f} ms | iters max {it.max()} mean {it.mean():.2f} | slowest instance: {ct[w,0]:.2f} ms, iters {it[w]}, qp_iters {qi[w]}, qp_fail {qf[w]}, set_qp {ct[w,1]:.2f} solve_qp {ct[w,2]:.2f} alpha {ct[w,3]:.2f} | median instance {np.median(ct[:,0]):.2f} ms")
    u = r["u0"]; x = mpc.sim_time_step(r["x0"], u)
