"""First on-GPU check: RobotData parity, one-cycle parity, and a rough timing of the named workload."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import mpcc_manipulator_b200 as M
from oracle import oracle as O

np.set_printoptions(precision=4, linewidth=200)
import __graft_entry__ as G
G.smoke()

nn = O.OracleNN()
ee, _, _ = O.fk(O.Q_HOME)
rng = np.random.default_rng(1)
# RobotData parity with an active obstacle on a few hundred samples
B, N = 64, 10
mpc = M.BatchMPC(B, N); mpc.setup_default(init_position=ee)
n = 300
q = O.Q_HOME + rng.uniform(-0.4, 0.4, (n, 7))
obs = np.c_[0.48 + rng.uniform(-0.1, 0.1, n), 0.218 + rng.uniform(-0.1, 0.1, n), 0.521 + rng.uniform(-0.1, 0.1, n), np.full(n, 5.0)]
rb = mpc.eval_robot_data(q, obs)
ref = np.stack([nn.robot_data(q[i], obs[i]) for i in range(n)])
names = dict(q=(0, 7), p=(7, 10), R=(10, 19), Jv=(19, 40), Jw=(40, 61), manip=(61, 62), dmanip=(62, 69), sel=(69, 70), dsel=(70, 77), obsr=(77, 78), env=(78, 87), denv=(87, 150))
for k, (a, b) in names.items():
    d = np.abs(rb[:, a:b] - ref[:, a:b]).max(); s = np.abs(ref[:, a:b]).max()
    print(f"  {k:7s} max abs diff {d:.3e}  rel {d / max(s, 1e-300):.3e}")
mpc.close()

# timing on the named workload
for (B, N) in [(4096, 20)]:
    mpc = M.BatchMPC(B, N); mpc.setup_default(init_position=ee)
    x0 = np.tile(np.r_[O.Q_HOME, 0., 0.], (B, 1)); x0[:, :7] += rng.uniform(-0.05, 0.05, (B, 7)); u0 = np.zeros((B, 8))
    for cyc in range(6):
        t0 = time.time(); r = mpc.run_cycle(x0, u0, want_horizon=False); dt = time.time() - t0
        st = mpc.stats()
        print(f"B={B} N={N} cycle {cyc}: {dt * 1e3:.2f} ms  solved={st['solved']} ok={st['ok']} sqp_iters={st['sqp_iters']} qp_iters={st['qp_iters']} qp_fail={st['qp_fail']}")
        u0 = r["u0"]; x0 = mpc.sim_time_step(r["x0"], u0)
    # kernel-level timing with events on the handle's stream
    dx = torch.from_numpy(x0).cuda(); du = torch.from_numpy(u0).cuda()
    s = torch.cuda.ExternalStream(mpc.stream)
    with torch.cuda.stream(s):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s); mpc.run_cycle_device(dx.data_ptr(), du.data_ptr()); e1.record(s)
    mpc.synchronize(); torch.cuda.synchronize()
    print(f"device-only cycle: {e0.elapsed_time(e1):.3f} ms")
    mpc.close()
