"""Per-instance SQP-loop durations of the named workload (diagnostic): distribution and relation to iteration counts."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mpcc_manipulator_b200 as M
from bench import synthetic_inputs, q_home

B, N = 4096, 20
mpc = M.BatchMPC(B, N); mpc.load_nn(); mpc.set_params(M.load_default_params())
ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee))
x0, u0 = synthetic_inputs(B, 0)
for c in range(14):
    r = mpc.run_cycle(x0, u0, want_horizon=False)
    tt = mpc.compute_time() * 1e3
    t = tt[:, 0]
    it = r["iters"]
    st = mpc.stats()
    print(f"cycle {c}: sqp ms min {t.min():.3f} p50 {np.median(t):.3f} p90 {np.percentile(t, 90):.3f} p99 {np.percentile(t, 99):.3f} max {t.max():.3f} | iters hist {np.bincount(np.minimum(it, 10))} | qp_iters {st['qp_iters']} qp_fail {st['qp_fail']}")
    for k in sorted(set(it.tolist())):
        m = it == k
        print(f"    iters={k}: n={m.sum()} mean {t[m].mean():.3f} ms max {t[m].max():.3f} | set_qp {tt[m,1].mean():.3f} solve_qp {tt[m,2].mean():.3f} get_alpha {tt[m,3].mean():.3f}")
    qi, qf = mpc.qp_counters()
    big = np.where(it >= 50)[0]
    for b in big[:4]:
        print(f"    straggler b={b}: iters {it[b]} qp_iters {qi[b]} qp_fail {qf[b]} status {r['status'][b]} x_in {np.array2string(x0[b], precision=4)} u_in {np.array2string(u0[b], precision=3)}")
    u0 = r["u0"]; x0 = mpc.sim_time_step(r["x0"], u0)
