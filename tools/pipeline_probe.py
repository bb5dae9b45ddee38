"""Experiment: G independent sub-batches (handles, streams) of the C2 workload running their closed loops concurrently."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mpcc_manipulator_b200 as M
from bench import synthetic_inputs, q_home, DevPtr

B, N, K = 4096, 20, 12
for G in (1, 2, 4, 8):
    Bs = B // G
    hs = []
    x0, u0 = synthetic_inputs(B, 0)
    for g in range(G):
        mpc = M.BatchMPC(Bs, N); mpc.load_nn(); mpc.set_params(M.load_default_params())
        ee = mpc.eval_robot_data(q_home()[None])[0, 7:10]
        mpc.set_tracks(M.load_track_json(None, ee))
        st = torch.cuda.ExternalStream(mpc.stream)
        with torch.cuda.stream(st):
            x = torch.from_numpy(x0[g * Bs:(g + 1) * Bs]).cuda(); xn = torch.empty_like(x); u = torch.from_numpy(u0[g * Bs:(g + 1) * Bs]).cuda()
            uo = torch.as_tensor(DevPtr(mpc.result_pointers()[0], (Bs, 8), "<f8"), device="cuda:0")
        hs.append([mpc, st, x, xn, u, uo])
    def step():
        for h in hs:
            mpc, st, x, xn, u, uo = h
            mpc.run_cycle_device(x.data_ptr(), u.data_ptr())
            with torch.cuda.stream(st):
                u.copy_(uo)
            mpc.sim_time_step_device(x.data_ptr(), u.data_ptr(), xn.data_ptr())
            h[2], h[3] = xn, x
    for _ in range(4): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K): step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    solved = sum(h[0].stats()["solved"] for h in hs)
    print(f"G={G}: {dt / K * 1e3:.2f} ms per step of {B} instances -> {B * K / dt:.0f} solves/s (solved last step {solved})")
    for h in hs: h[0].close()
