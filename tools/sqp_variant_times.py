"""Kernel times of the C2 closed loop for a given build of the library (MPCC_B200_LIB=<path>): mean of k_sqp_warp / k_mlp over the
bench window (cycles 3..22 after a cold start) and in steady state (cycles 40..59).  Uses only the round-1 ABI so that older
builds can be compared."""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import mpcc_manipulator_b200 as M
B, N = 4096, 20
q_home = np.array([0, 0, 0, -np.pi / 2, 0, np.pi / 2, np.pi / 4])
rng = np.random.default_rng(0)
x_host = np.tile(np.r_[q_home, 0.0, 0.0], (B, 1)); x_host[:, :7] += rng.uniform(-0.05, 0.05, (B, 7))
mpc = M.BatchMPC(B, N); mpc.load_nn(); mpc.set_params(M.load_default_params())
ee = mpc.eval_robot_data(q_home[None])[0, 7:10]
mpc.set_tracks(M.load_track_json(None, ee))
stream = torch.cuda.ExternalStream(mpc.stream)
p_u = mpc.result_pointers()[0]
class DevPtr:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}
with torch.cuda.stream(stream):
    x = torch.from_numpy(x_host).cuda(); xn = torch.empty_like(x); u = torch.zeros((B, 8), dtype=torch.float64, device="cuda")
    u_out = torch.as_tensor(DevPtr(p_u, (B, 8), "<f8"), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
mpc.set_profiling(True)
kt = []
for c in range(60):
    with torch.cuda.stream(stream):
        flush.zero_()
    mpc.run_cycle_device(x.data_ptr(), u.data_ptr())
    with torch.cuda.stream(stream):
        u.copy_(u_out)
    mpc.sim_time_step_device(x.data_ptr(), u.data_ptr(), xn.data_ptr())
    x, xn = xn, x
    kt.append(mpc.kernel_times())
kt = np.array(kt)
st = mpc.stats()
print(os.environ.get("MPCC_B200_LIB", "default"), "| bench window (3..22): k_mlp %.3f k_sqp %.3f | steady (40..59): k_mlp %.3f k_sqp %.3f | per-cycle sqp:" % (kt[3:23, 2].mean(), kt[3:23, 3].mean(), kt[40:, 2].mean(), kt[40:, 3].mean()),
      " ".join("%.1f" % v for v in kt[:30, 3]), "| last", st)
