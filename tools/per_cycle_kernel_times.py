"""Per-cycle kernel times (CUDA events of the library) of the bench workload: k_prologue, k_kin, k_mlp, k_sqp_warp."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import mpcc_manipulator_b200 as M
from bench import synthetic_inputs, q_home
B, N = 4096, 20
import os
mpc = M.BatchMPC(B, N, flags=int(os.environ.get('MPCC_BENCH_FLAGS', '0'))); mpc.load_nn(); mpc.set_params(M.load_default_params())
ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee))
mpc.set_profiling(True)
x0, u0 = synthetic_inputs(B, 0)
tot = []
for c in range(23):
    r = mpc.run_cycle(x0, u0, want_horizon=False)
    kt = mpc.kernel_times()
    it = r["iters"]
    tot.append(kt[3])
    print(f"cycle {c:2d}: k_mlp {kt[2]:6.2f}  k_sqp {kt[3]:6.2f} ms | max iters {it.max():3d}  n(2 iters) {(it == 2).sum():4d}  stragglers {(it >= 50).sum()}")
    u0 = r["u0"]; x0 = mpc.sim_time_step(r["x0"], u0)
print("mean k_sqp over cycles 3..22:", np.mean(tot[3:23]))
