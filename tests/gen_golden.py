"""Generates tests/golden/*.npz from the CPU oracle (oracle/liborc.so).

The reference itself cannot be built or imported in the build container (Eigen, RBDL, OSQP, OsqpEigen absent;
SURVEY.md section 8c), and its own tests hold no golden vector for this path, so these fixtures are outputs of the
restated oracle after it passed tests/test_oracle_pins.py.  They freeze the oracle (regression pin) and give the
GPU tier a reference that does not need the oracle to be rebuilt identically.  Run:  python tests/gen_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import oracle as O  # noqa: E402
from helpers import make_horizon  # noqa: E402

OUT = ROOT / "tests" / "golden"


def main():
    OUT.mkdir(exist_ok=True)
    nn = O.OracleNN()
    ee = O.fk(O.Q_HOME)[0]
    X, Y, Z, R = O.load_track()
    X, Y, Z = O.shift_track(X, Y, Z, ee)
    p = O.load_params()
    Ts = p["Ts"]

    # 1. RobotData (kinematics + both networks), dummy and active obstacle
    rng = np.random.default_rng(100)
    n = 96
    q = O.Q_HOME + rng.uniform(-0.6, 0.6, (n, 7))
    obs = np.c_[0.48 + rng.uniform(-0.15, 0.15, n), 0.218 + rng.uniform(-0.15, 0.15, n), 0.521 + rng.uniform(-0.15, 0.15, n), np.full(n, 5.0)]
    obs[: n // 3] = [3., 3., 3., 0.]
    rb = np.stack([nn.robot_data(q[i], obs[i]) for i in range(n)])
    np.savez_compressed(OUT / "robot_data.npz", q=q, obs=obs, rb=rb)

    # 2. track evaluation on the default (shifted) track
    o = O.OracleMPC(N=10, nn=nn); o.set_track(X, Y, Z, R)
    L = o.track_length
    s = np.r_[0.0, L, L + 0.3, -0.2, rng.uniform(0, L, 60)]
    ev = []
    for si in s:
        r = o.track_eval(si)
        ev.append(np.r_[r["pos"], r["dpos"], r["ddpos"], r["R"].ravel(), r["dR"]])
    np.savez_compressed(OUT / "track_eval.npz", s=s, out=np.array(ev), length=L)

    # 3. flat QP of one linearisation (N = 10), with an active obstacle and dynamics defects
    N = 10
    hor = make_horizon(O, rng, N, Ts)
    hor[1:, :9] += rng.normal(0, 1e-3, (N, 9))
    ob = (0.48, 0.218, 0.521, 5.0)
    rbh = np.stack([nn.robot_data(hor[k, :7], ob) for k in range(N + 1)])
    cur_u = np.r_[rng.uniform(-0.1, 0.1, 7), 0.0]
    qp = o.build_qp(hor, rbh, cur_u)
    ok, z, it = O.solve_qp_dense(qp["P"], qp["q"], qp["A"], qp["l"] - qp["c"], qp["u"] - qp["c"])
    assert ok
    np.savez_compressed(OUT / "flat_qp_n10.npz", hor=hor, rb=rbh, cur_u=cur_u, P=qp["P"], q=qp["q"], A=qp["A"], l=qp["l"], u=qp["u"], c=qp["c"], obj=qp["obj"], z=z)

    # 4. configuration C1: closed loop from q_home, N = 10 (main.cpp:57-114), first 40 cycles
    x = np.r_[O.Q_HOME, 0., 0.]; u = np.zeros(8)
    o = O.OracleMPC(N=10, nn=nn); o.set_track(X, Y, Z, R)
    rec = dict(x_in=[], u_in=[], x_out=[], u_out=[], status=[], iters=[], margin=[], hor=[])
    for c in range(40):
        r = o.run(x, u)
        rec["x_in"].append(x); rec["u_in"].append(u); rec["x_out"].append(r["x0"]); rec["u_out"].append(r["u0"])
        rec["status"].append(r["status"]); rec["iters"].append(r["iters"]); rec["margin"].append(o.last_filter_margin()); rec["hor"].append(r["horizon"])
        u = r["u0"]; x = O.sim_time_step(r["x0"], u, Ts)
    np.savez_compressed(OUT / "closed_loop_c1.npz", **{k: np.array(v) for k, v in rec.items()})

    # 5. configuration C2 in small: 12 perturbed starts, N = 20, 3 closed-loop cycles each
    rng = np.random.default_rng(0)
    B, N = 12, 20
    q0 = O.Q_HOME + rng.uniform(-0.05, 0.05, (B, 7))
    rec = dict(x_in=np.zeros((3, B, 9)), u_in=np.zeros((3, B, 8)), x_out=np.zeros((3, B, 9)), u_out=np.zeros((3, B, 8)),
               status=np.zeros((3, B), int), iters=np.zeros((3, B), int), margin=np.zeros((3, B)))
    for b in range(B):
        o = O.OracleMPC(N=N, nn=nn); o.set_track(X, Y, Z, R)
        x = np.r_[q0[b], 0., 0.]; u = np.zeros(8)
        for c in range(3):
            r = o.run(x, u)
            rec["x_in"][c, b] = x; rec["u_in"][c, b] = u; rec["x_out"][c, b] = r["x0"]; rec["u_out"][c, b] = r["u0"]
            rec["status"][c, b] = r["status"]; rec["iters"][c, b] = r["iters"]; rec["margin"][c, b] = o.last_filter_margin()
            u = r["u0"]; x = O.sim_time_step(r["x0"], u, Ts)
    np.savez_compressed(OUT / "batch_c2_small.npz", **rec)

    # 6. configuration C3 in small: active moving obstacle, tightened model parameters (python/main_w_sim.py:21-46)
    rng = np.random.default_rng(1)
    B, N = 8, 20
    pc3 = O.load_params(overrides={"model": {"tol_envcol": 1.0, "tol_sing": 0.018, "desired_ee_velocity": 0.1}})
    q0 = O.Q_HOME + rng.uniform(-0.05, 0.05, (B, 7))
    obs = np.c_[np.array([0.48, 0.218, 0.521]) + rng.uniform(-0.05, 0.05, (B, 3)), np.full(B, 5.0)]
    rec = dict(x_in=np.zeros((3, B, 9)), u_in=np.zeros((3, B, 8)), obs=np.zeros((3, B, 4)), x_out=np.zeros((3, B, 9)), u_out=np.zeros((3, B, 8)),
               status=np.zeros((3, B), int), iters=np.zeros((3, B), int), margin=np.zeros((3, B)))
    for b in range(B):
        o = O.OracleMPC(N=N, nn=nn, params=pc3); o.set_track(X, Y, Z, R)
        x = np.r_[q0[b], 0., 0.]; u = np.zeros(8); ob = obs[b].copy()
        for c in range(3):
            r = o.run(x, u, ob)
            rec["x_in"][c, b] = x; rec["u_in"][c, b] = u; rec["obs"][c, b] = ob; rec["x_out"][c, b] = r["x0"]; rec["u_out"][c, b] = r["u0"]
            rec["status"][c, b] = r["status"]; rec["iters"][c, b] = r["iters"]; rec["margin"][c, b] = o.last_filter_margin()
            u = r["u0"]; x = O.sim_time_step(r["x0"], u, Ts); ob[2] += 0.05 * Ts
    np.savez_compressed(OUT / "batch_c3_small.npz", **rec)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
