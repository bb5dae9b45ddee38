"""Latency mode (C5: 64 instances, N = 40, eps_prim = 0.01): distribution of the end-to-end cycle time and what sets it --
per-kernel event times, per-instance SQP times, SQP / interior-point iteration counts."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mpcc_manipulator_b200 as M
from bench import q_home
B, N = 64, 40
CYCLES = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(3)
FLAGS = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # 2: warp-per-instance kernel, 4: CTA-per-instance kernel (default for this batch)
mpc = M.BatchMPC(B, N, flags=FLAGS); mpc.load_nn(); mpc.set_params(M.load_default_params(overrides={"sqp.eps_prim": 0.01}))
ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee)); mpc.set_profiling(True)
x = np.tile(np.r_[q_home(), 0., 0.], (B, 1)); x[:, :7] += rng.uniform(-0.05, 0.05, (B, 7)); u = np.zeros((B, 8))
rows = []
for c in range(CYCLES):
    t0 = time.perf_counter(); r = mpc.run_cycle(x, u, want_horizon=False); dt = (time.perf_counter() - t0) * 1e3
    kt = mpc.kernel_times(); ct = mpc.compute_time() * 1e3; qi, qf = mpc.qp_counters(); it = r["iters"]
    w = int(np.argmax(ct[:, 0]))
    rows.append((dt, kt[2], kt[3], it.max(), it.mean(), ct[w, 0], it[w], qi[w], qf[w], np.median(ct[:, 0]), qi.max(), (r["status"] != 0).sum()))
    u = r["u0"]; x = mpc.sim_time_step(r["x0"], u)
    if len(sys.argv) > 3:
        print("cycle", c, "max iters", it.max(), "qp_iters max", qi.max(), "fails", qf.sum(), "non-solved", (r["status"] != 0).sum(), flush=True)
a = np.array(rows[20:], dtype=float)
print("cycles", len(a))
print("e2e ms  p50 %.2f p90 %.2f p99 %.2f max %.2f" % tuple(np.percentile(a[:, 0], [50, 90, 99, 100])))
print("k_mlp ms p50 %.2f | k_sqp ms p50 %.2f p90 %.2f p99 %.2f" % (np.median(a[:, 1]), *np.percentile(a[:, 2], [50, 90, 99])))
print("max SQP iters per cycle: p50 %d p90 %d p99 %d max %d; mean iters %.2f" % (*np.percentile(a[:, 3], [50, 90, 99, 100]), a[:, 4].mean()))
print("slowest instance ms: p50 %.2f p90 %.2f p99 %.2f ; median instance ms p50 %.2f" % (*np.percentile(a[:, 5], [50, 90, 99]), np.median(a[:, 9])))
sel = a[:, 7] > 0
print("us per interior-point iteration of the slowest instance (incl. linearisations): p50 %.1f" % np.median(1e3 * a[sel, 5] / a[sel, 7]))
print("cycles with a non-SOLVED instance:", int((a[:, 11] > 0).sum()))
for thr in (3, 5, 10, 20, 50):
    print("cycles with e2e > %d ms: %d" % (thr, int((a[:, 0] > thr).sum())))
idx = np.argsort(-a[:, 0])[:8]
for i in idx:
    print("  slow cycle %d: e2e %.2f ms, slowest instance %.2f ms iters %d qp_iters %d qp_fail %d" % (i + 20, a[i, 0], a[i, 5], a[i, 6], a[i, 7], a[i, 8]))
print(json.dumps({"p50_ms": float(np.percentile(a[:, 0], 50)), "p99_ms": float(np.percentile(a[:, 0], 99)), "max_ms": float(a[:, 0].max())}))
