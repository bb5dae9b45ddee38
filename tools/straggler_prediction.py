import sys; sys.path.insert(0, '/root/repo')
import numpy as np
import mpcc_manipulator_b200 as M
from bench import synthetic_inputs, q_home
B, N = 4096, 20
mpc = M.BatchMPC(B, N); mpc.load_nn(); mpc.set_params(M.load_default_params())
ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee))
x0, u0 = synthetic_inputs(B, 0)
hist = []
for c in range(24):
    r = mpc.run_cycle(x0, u0, want_horizon=False)
    it = r["iters"].copy()
    t = mpc.compute_time()[:, 0] * 1e3
    s = np.where(it >= 50)[0]
    prev = np.max(np.stack(hist[-4:]), axis=0) if hist else np.zeros(B, int)
    pred = prev >= 15
    print(f"cycle {c}: stragglers {len(s)}, of which predicted {int(pred[s].sum())}; predicted set size {int(pred.sum())}; straggler ms {np.round(t[s], 1)} predicted? {pred[s].astype(int)}")
    if len(s) and hist:
        print("      their iteration counts in the previous cycles:", [[int(h[i]) for h in hist[-3:]] for i in s], "| population share with >= 3 iterations last cycle:", float((hist[-1] >= 3).mean()))
    hist.append(it)
    u0 = r["u0"]; x0 = mpc.sim_time_step(r["x0"], u0)
