"""Small closed-loop run for compute-sanitizer (memcheck / racecheck): B = 6, N = 10, 3 cycles, active obstacle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mpcc_manipulator_b200 as M
from bench import q_home
B, N = 6, 10
mpc = M.BatchMPC(B, N); mpc.load_nn(); mpc.set_params(M.load_default_params())
ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee))
rng = np.random.default_rng(0)
x = np.tile(np.r_[q_home(), 0., 0.], (B, 1)); x[:, :7] += rng.uniform(-0.05, 0.05, (B, 7)); u = np.zeros((B, 8))
obs = np.tile([0.48, 0.218, 0.521, 5.0], (B, 1))
for c in range(3):
    r = mpc.run_cycle(x, u, obs)
    u = r["u0"]; x = mpc.sim_time_step(r["x0"], u)
print("ok", r["status"].tolist(), r["iters"].tolist())
mpc.close()
