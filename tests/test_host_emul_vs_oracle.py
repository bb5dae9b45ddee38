"""CPU tier: the product's device functions (csrc/dev_*.cuh), compiled for the host by tests/emul, against the
oracle.  These are the same inline functions the CUDA kernels call; the GPU tier repeats the comparisons on
the device through the C ABI.  Tolerances: linearisations 1e-9 relative (north_star), QP/SQP steps within the
reference QP solver's tolerance (OSQP eps_abs = 1e-4, osqp_interface.cpp:623)."""
from pathlib import Path

import numpy as np
import pytest

from helpers import Emul, f64, flat_params, assemble_flat_qp_from_lin, make_horizon, step_to_flat, unpack_sym9
import helpers as H

REL = 1e-9


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def emu():
    return Emul()


@pytest.fixture(scope="module")
def setup(O, nn, track_wp, emu):
    p = O.load_params()
    pf = flat_params(p)
    table = emu.fit_track(*track_wp)
    return p, pf, table


def test_kinematics(O, emu, rng):
    for _ in range(50):
        q = O.Q_HOME + rng.uniform(-1.0, 1.0, 7)
        k = emu.kin(q)
        p, R, J = O.fk(q)
        m, dm = O.manip(q)
        assert rel_err(k["p"], p) < REL and rel_err(k["R"], R) < REL
        assert rel_err(k["Jv"], J[:3]) < REL and rel_err(k["Jw"], J[3:]) < REL
        assert abs(k["manip"] - m) < REL * abs(m)
        # central differences with delta 1e-4 amplify rounding by ~1e4: 1e-9 relative still holds
        assert rel_err(k["dmanip"], dm) < REL


def test_track_fit_and_eval(O, nn, emu, track_wp, setup, rng):
    p, pf, table = setup
    o = O.OracleMPC(N=10, nn=nn)
    o.set_track(*track_wp)
    s_o, X_o, Y_o, Z_o, R_o = o.track_table()
    assert rel_err(table[:100], s_o) < REL
    assert abs(table[100 * (1 + 12 + 9 + 3 + 2) + 1] - o.track_length) < 1e-12  # TrackTable::length
    L = o.track_length
    for s in np.r_[0.0, L, L + 0.5, -0.1, rng.uniform(0, L, 100), s_o[1:99], s_o[1:99] - 1e-13]:
        e = emu.track_eval(table, s)
        r = o.track_eval(s)
        ref = np.r_[r["pos"], r["dpos"], r["ddpos"], r["R"].ravel(), r["dR"]]
        assert np.abs(e - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), s


def test_track_heterogeneous_family(O, nn, emu, rng):
    # cpp/Params/track.py family (config 4): random a, b, c
    o = O.OracleMPC(N=10, nn=nn)
    for _ in range(5):
        a, b = rng.uniform(1.5, 3, 2); c = rng.uniform(0, 2.5)
        t = np.linspace(np.pi / 2, 5 * np.pi / 2, 100); r = 0.1
        X, Y, Z = a * r * np.sin(t) + 0.3, b * r * np.sin(2 * t), c * r * np.cos(t) + 0.5
        R = np.tile(np.diag([1., -1., -1.]).ravel(), (100, 1))
        table = emu.fit_track(X, Y, Z, R)
        o.set_track(X, Y, Z, R)
        for s in rng.uniform(0, o.track_length, 20):
            r_ = o.track_eval(s)
            ref = np.r_[r_["pos"], r_["dpos"], r_["ddpos"], r_["R"].ravel(), r_["dR"]]
            assert np.abs(emu.track_eval(table, s) - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_rotating_track_orientation_spline(O, nn, emu, rng):
    # an orientation that actually rotates along the path exercises Log/Exp and the rotation-spline derivative
    o = O.OracleMPC(N=10, nn=nn)
    t = np.linspace(0, 1, 100)
    X, Y, Z = 0.4 + 0.2 * t, 0.3 * np.sin(2 * t), 0.5 + 0.1 * t
    R = []
    for a in t:
        ca, sa = np.cos(0.8 * a), np.sin(0.8 * a)
        Rz = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]])
        cb, sb = np.cos(0.5 * a), np.sin(0.5 * a)
        Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
        R.append((Rz @ Ry @ np.diag([1., -1., -1.])).ravel())
    R = np.array(R)
    table = emu.fit_track(X, Y, Z, R)
    o.set_track(X, Y, Z, R)
    for s in rng.uniform(0, o.track_length, 50):
        r_ = o.track_eval(s)
        ref = np.r_[r_["pos"], r_["dpos"], r_["ddpos"], r_["R"].ravel(), r_["dR"]]
        assert np.abs(emu.track_eval(table, s) - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        assert np.abs(r_["dR"]).max() > 1e-3


def test_projection(O, nn, emu, track_wp, setup, rng):
    p, pf, table = setup
    o = O.OracleMPC(N=10, nn=nn)
    o.set_track(*track_wp)
    L = o.track_length
    for _ in range(200):
        s = rng.uniform(0, L)
        ee = o.track_eval(min(max(s + rng.normal(0, 0.05), 0), L))["pos"] + rng.normal(0, 0.01, 3)
        assert abs(emu.project(table, p["model"][0], s, ee) - o.project(s, ee)) < 1e-12
    # far away: global nearest knot branch (arc_length_spline.cpp:326-352)
    ee = o.track_eval(1.5)["pos"] + 0.001
    assert abs(emu.project(table, p["model"][0], 0.1, ee) - o.project(0.1, ee)) < 1e-12


def _stage_inputs(O, nn, rng, N, Ts, obs=None):
    hor = make_horizon(O, rng, N, Ts)
    rb = np.stack([nn.robot_data(hor[k, :7], obs if obs is not None else (3., 3., 3., 0.)) for k in range(N + 1)])
    cur_u = np.r_[rng.uniform(-0.1, 0.1, 7), 0.0]
    return hor, rb, cur_u


@pytest.mark.parametrize("N", [10, 20])
def test_stage_linearisation_vs_flat_qp(O, nn, emu, track_wp, setup, rng, N):
    """Per-stage blocks (Cost::getCost, Constraints::getConstraints, Bounds, dynamics) against the oracle's dense
    osqp_interface-style assembly, entry by entry at 1e-9 relative."""
    p, pf, table = setup
    Ts = p["Ts"]
    o = O.OracleMPC(N=N, nn=nn)
    o.set_track(*track_wp)
    Tx, Tu = p["norm"][:9], p["norm"][9:]
    for trial in range(3):
        obs = None if trial == 0 else (0.48, 0.218, 0.521, 5.0)
        hor, rb, cur_u = _stage_inputs(O, nn, rng, N, Ts, obs)
        if trial == 2:
            hor[1:, :9] += rng.normal(0, 1e-3, (N, 9))  # dynamics defects
        lin = []
        for k in range(N + 1):
            up = cur_u[:7] if k == 0 else hor[k - 1, 9:16]
            un = hor[k + 1, 9:16] if k < N else np.zeros(7)
            xn = hor[k + 1, :9] if k < N else np.zeros(9)
            lin.append(emu.stage_lin(pf, table, Ts, N, k, hor[k, :9], hor[k, 9:], up, un, xn, rb[k]))
        lin = np.array(lin)
        ref = o.build_qp(hor, rb, cur_u)
        P, q = assemble_flat_qp_from_lin(lin, pf, N, Ts)
        assert rel_err(P, ref["P"]) < REL
        assert rel_err(q, ref["q"]) < REL
        assert abs(lin[:, H.LOBJ].sum() - ref["obj"]) < REL * abs(ref["obj"])
        nx = 9 * (N + 1)
        A, l, u, c = ref["A"], ref["l"], ref["u"], ref["c"]
        lo, hi = l - c, u - c
        E0, B0 = 0, nx
        U0 = B0 + nx; D0 = U0 + 8 * N; P0 = D0 + 8 * N
        gap_ref = np.sum(np.maximum(l - c, 0) + np.maximum(c - u, 0))
        assert abs(lin[:, H.LGAP].sum() - gap_ref) <= 1e-9 * max(1.0, gap_ref)
        for k in range(N + 1):
            L = lin[k]
            if k >= 1:  # dynamics rows: xi_k - A xi_{k-1} - B nu_{k-1} = -c  (b of stage k-1)
                assert np.allclose(lin[k - 1][H.Lb:H.Lb + 9], lo[E0 + 9 * k:E0 + 9 * k + 9], rtol=0, atol=1e-9 * max(1, np.abs(lo).max()) * 0 + 1e-12)
            # state box rows (pre-quirk): Tx * xi in [lx - x, ux - x]
            assert np.allclose(L[H.LXLO:H.LXLO + 9] * Tx, lo[B0 + 9 * k:B0 + 9 * k + 9], rtol=1e-9, atol=1e-12)
            assert np.allclose(L[H.LXHI:H.LXHI + 9] * Tx, hi[B0 + 9 * k:B0 + 9 * k + 9], rtol=1e-9, atol=1e-12)
            if k < N:
                # rate rows: (Tu/Ts)(nu_k - nu_{k-1}) in [.,.]
                assert np.allclose(L[H.LDLO:H.LDLO + 7] * Tu[:7] / Ts, lo[D0 + 8 * k:D0 + 8 * k + 7], rtol=1e-9, atol=1e-10)
                assert np.allclose(L[H.LDHI:H.LDHI + 7] * Tu[:7] / Ts, hi[D0 + 8 * k:D0 + 8 * k + 7], rtol=1e-9, atol=1e-10)
                # polytopic rows
                pg = L[H.LPG:H.LPG + 77].reshape(11, 7); pd = L[H.LPD:H.LPD + 11]; prhs = L[H.LPRHS:H.LPRHS + 11]
                rows = A[P0 + 11 * k:P0 + 11 * k + 11]
                ax = pd[:, None] * pg * Tx[None, :7]
                au = -pg * Tu[None, :7]
                assert rel_err(ax, rows[:, 9 * k:9 * k + 7]) < REL
                assert rel_err(au, rows[:, nx + 8 * k:nx + 8 * k + 7]) < REL
                assert np.allclose(prhs, hi[P0 + 11 * k:P0 + 11 * k + 11], rtol=1e-9, atol=1e-12)
        # dynamics matrix entries of the oracle are the closed form the product hard-codes
        k = 2
        blk = A[E0 + 9 * k:E0 + 9 * k + 9]
        assert np.allclose(blk[:, 9 * k:9 * k + 9], np.eye(9))
        Ad = -blk[:, 9 * (k - 1):9 * k]
        exp = np.eye(9); exp[7, 8] = Ts * Tx[8] / Tx[7]
        assert np.allclose(Ad, exp, atol=1e-15)
        Bd = -blk[:, nx + 8 * (k - 1):nx + 8 * k]
        expB = np.zeros((9, 8)); expB[:7, :7] = np.diag(Ts * Tu[:7] / Tx[:7]); expB[7, 7] = 0.5 * Ts * Ts * Tu[7] / Tx[7]; expB[8, 7] = Ts * Tu[7] / Tx[8]
        assert np.allclose(Bd, expB, atol=1e-15)


@pytest.mark.parametrize("N", [10, 20])
def test_structured_qp_vs_dense_oracle_qp(O, nn, emu, track_wp, setup, rng, N):
    """The stage-structured interior-point/Riccati QP solve against the oracle's generic dense QP solver on the
    reference's flat QP (quirk 1 rows included): primal optimum within 1e-6 (reference tolerance: 1e-4)."""
    p, pf, table = setup
    Ts = p["Ts"]
    o = O.OracleMPC(N=N, nn=nn)
    o.set_track(*track_wp)
    for trial in range(3):
        obs = None if trial == 0 else (0.48, 0.218, 0.521, 5.0)
        hor, rb, cur_u = _stage_inputs(O, nn, rng, N, Ts, obs)
        ok, step, iters, res = emu.solve_qp(pf, table, Ts, N, hor, rb, cur_u)
        assert ok, res
        ref = o.build_qp(hor, rb, cur_u)
        ok2, z, it2 = O.solve_qp_dense(ref["P"], ref["q"], ref["A"], ref["l"] - ref["c"], ref["u"] - ref["c"])
        assert ok2
        zs = step_to_flat(step, N)
        # The Gauss-Newton Hessian is only regularised by 1e-6 I (cost.cpp:353), so the minimiser is weakly determined
        # along some directions: two solvers that both meet 1e-9 KKT residuals can differ by ~1e-5 in z.  The bar is
        # the reference QP solver's own tolerance (OSQP eps_abs = 1e-4), and an objective that matches to 1e-7 relative (both IPMs stop at a 1e-9 gap per constraint).
        fz = lambda v: 0.5 * v @ ref["P"] @ v + ref["q"] @ v
        assert np.abs(zs - z).max() < 1e-4, (np.abs(zs - z).max(), iters, it2)
        assert abs(fz(zs) - fz(z)) < 1e-7 * (1 + abs(fz(z))), (fz(zs), fz(z))
        # KKT sanity on the flat problem: feasibility of the structured solution
        Az = ref["A"] @ zs
        assert np.all(Az >= ref["l"] - ref["c"] - 1e-7) and np.all(Az <= ref["u"] - ref["c"] + 1e-7)


@pytest.mark.parametrize("N", [10, 20])
def test_sqp_loop_vs_oracle(O, nn, emu, track_wp, setup, rng, N):
    """solveOCP (osqp_interface.cpp:398-590): same iteration count, alphas, status; steps within the QP tolerance."""
    p, pf, table = setup
    Ts = p["Ts"]
    o = O.OracleMPC(N=N, nn=nn)
    o.set_track(*track_wp)
    ties = 0
    for trial in range(8):
        q0 = O.Q_HOME + rng.uniform(-0.05, 0.05, 7)
        x0 = np.r_[q0, 0.0, 0.0]
        hor = np.tile(np.r_[x0, np.zeros(8)], (N + 1, 1))  # generateNewInitialGuess
        rb = np.stack([nn.robot_data(hor[k, :7]) for k in range(N + 1)])
        cur_u = np.zeros(8)
        a = emu.solve_ocp(pf, table, Ts, N, hor, rb, cur_u)
        b = o.solve_ocp(hor, rb, cur_u)
        assert a["status"] == b["status"] == 0
        # iteration by iteration while the line-search decisions agree: QP steps within the reference QP tolerance
        n_same = 0
        for i in range(min(len(a["alphas"]), len(b["alphas"]))):
            assert np.abs(step_to_flat(a["steps"][i], N) - b["steps"][i]).max() < 1e-4
            if a["alphas"][i] != b["alphas"][i]:
                break
            n_same += 1
        assert n_same >= 1  # the first iteration has an empty filter: always alpha = 1
        if n_same == len(b["alphas"]) == len(a["alphas"]):
            assert a["iters"] == b["iters"]
            assert np.abs(a["horizon"] - b["horizon"]).max() < 1e-4
        else:
            # A differing accept/reject decision is only legitimate when the filter comparison hinges on solver noise:
            # after a full step both l1 violations are ~1e-9 numbers, and "gap >= filter.gap" (osqp_interface.cpp:780)
            # compares them.  The oracle reports the robustness margin of its own comparisons.
            assert o.last_filter_margin() < 1e-6, o.last_filter_margin()
            ties += 1
    assert ties <= 6


def test_prologue_vs_oracle_cycles(O, nn, emu, track_wp, setup, rng):
    """runMPC_ prologue (mpc.cpp:104-124): projection, vs estimate, warm-start shift / regeneration, followed over
    several closed-loop cycles of the oracle."""
    p, pf, table = setup
    Ts, N = p["Ts"], 10
    o = O.OracleMPC(N=N, nn=nn)
    o.set_track(*track_wp)
    x = np.r_[O.Q_HOME + rng.uniform(-0.02, 0.02, 7), 0.0, 0.0]
    u = np.zeros(8)
    warm = np.zeros((N + 1, 17)); valid, failed = 0, 0
    for cyc in range(4):
        xe, we, valid_e, failed_e = emu.prologue(pf, table, Ts, N, x, u, warm, valid, failed)
        r = o.run(x, u)
        assert abs(xe[7] - r["x0"][7]) < 1e-12 and abs(xe[8] - r["x0"][8]) < 1e-12
        if cyc == 0:
            assert np.allclose(we, np.tile(np.r_[xe, np.zeros(8)], (N + 1, 1)))
        warm, valid, failed = o.warm_state()
        u = r["u0"]
        x = O.sim_time_step(r["x0"], u, Ts)
    # shifted guess: product prologue on the oracle's persistent warm start equals the oracle's next initial guess
    xe, we, v2, f2 = emu.prologue(pf, table, Ts, N, x, u, warm, valid, failed)
    assert v2 == 1
    assert np.allclose(we[0, :9], xe) and np.allclose(we[1:N - 1], warm[2:N]) and np.allclose(we[N - 1], we[N - 2])
    assert np.allclose(we[N, :9], O.rk4(we[N - 1, :9], we[N - 1, 9:], Ts)) or we[N, 7] == o.track_length


# ---- the warp-per-instance formulation (csrc/sqp_warp.cuh), executed phase by phase on the host ----------------
@pytest.mark.parametrize("N", [10, 20, 40])
def test_warp_qp_matches_thread_formulation_and_lane_order(O, nn, emu, track_wp, setup, rng, N):
    """Same interior-point iteration, different work distribution: steps agree to rounding, and the result does not
    depend on the order in which the 32 lanes of a phase are executed (no intra-phase data dependence)."""
    p, pf, table = setup
    Ts = p["Ts"]
    for trial in range(3):
        obs = None if trial == 0 else (0.48, 0.218, 0.521, 5.0)
        hor, rb, cur_u = _stage_inputs(O, nn, rng, N, Ts, obs)
        ok1, s1, it1, r1 = emu.solve_qp(pf, table, Ts, N, hor, rb, cur_u)
        ok2, s2, it2, r2 = emu.warp_solve_qp(pf, table, Ts, N, hor, rb, cur_u)
        ok3, s3, it3, r3 = emu.warp_solve_qp(pf, table, Ts, N, hor, rb, cur_u, reverse=True)
        assert ok1 == ok2 == ok3
        if ok1:
            assert abs(it1 - it2) <= 1 and it2 == it3
            # same iteration count: the same iterates up to rounding; one interior-point iteration apart (the tail is superlinear, the stopping test
            # can fall on either side): both are solutions to the 1e-9 KKT tolerance and differ by the termination slack along weakly convex directions
            assert np.abs(s1 - s2).max() < (1e-7 if it1 == it2 else 2e-5) and np.abs(s2 - s3).max() < 1e-9


@pytest.mark.parametrize("N", [10, 20])
def test_warp_sqp_loop_vs_oracle(O, nn, emu, track_wp, setup, rng, N):
    p, pf, table = setup
    Ts = p["Ts"]
    o = O.OracleMPC(N=N, nn=nn)
    o.set_track(*track_wp)
    for trial in range(6):
        q0 = O.Q_HOME + rng.uniform(-0.05, 0.05, 7)
        hor = np.tile(np.r_[q0, 0.0, 0.0, np.zeros(8)], (N + 1, 1))
        rb = np.stack([nn.robot_data(hor[k, :7]) for k in range(N + 1)])
        cur_u = np.zeros(8)
        a = emu.warp_solve_ocp(pf, table, Ts, N, hor, rb, cur_u, reverse=bool(trial % 2))
        o.set_forced_decisions([int(x == 1.0) for x in a["alphas"]])
        b = o.solve_ocp(hor, rb, cur_u)
        nat, mg = o.decision_log()
        o.set_forced_decisions([])
        assert a["status"] == b["status"] == 0 and a["iters"] == b["iters"]
        assert np.allclose(a["alphas"], b["alphas"])
        for i in range(len(b["steps"])):
            assert np.abs(step_to_flat(a["steps"][i], N) - b["steps"][i]).max() < 1e-4
            assert nat[i] == int(a["alphas"][i] == 1.0) or mg[i] < 1e-6
        assert a["accept_mask"] == sum(int(x == 1.0) << i for i, x in enumerate(a["alphas"]))


def test_primal_infeasibility_certificate(O, nn, track_wp):
    """A QP of the latency configuration (N = 40, eps_prim = 0.01; instance 3, cycle 13 of the C5 closed loop) whose
    constraints are inconsistent by 3e-4: the reference's OSQP would report PrimalInfeasible (osqp_interface.cpp:495-497);
    the interior point used to burn all 60 iterations on it.  The Farkas test on the diverging multipliers stops it early.
    Independent confirmation: an LP phase 1 (scipy HiGHS) on the ORACLE's flat constraint matrix has a positive optimum,
    and the oracle's dense solver fails on the same QP."""
    from scipy.optimize import linprog
    g = np.load(Path(__file__).resolve().parent / "golden" / "infeasible_qp_n40.npz")
    N = 40
    emu = Emul()
    p = O.load_params(overrides={"sqp": {"eps_prim": 0.01}}); pf = flat_params(p)
    table = emu.fit_track(*track_wp)
    for rev in (False, True):
        ok, step, it, res = emu.warp_solve_qp(pf, table, p["Ts"], N, g["warm"], g["rb"], g["u"], reverse=rev)
        assert not ok and it < 30, (ok, it)
    o = O.OracleMPC(N=N, nn=nn, params=p); o.set_track(*track_wp)
    qp = o.build_qp(g["warm"], g["rb"], g["u"])
    A, lo, hi = qp["A"], qp["l"] - qp["c"], qp["u"] - qp["c"]
    okd, _, _ = O.solve_qp_dense(qp["P"], qp["q"], A, lo, hi)
    assert not okd
    keep = np.abs(A).sum(1) > 0
    A, lo, hi = A[keep], lo[keep], hi[keep]
    n = A.shape[1]
    fu, fl = hi < 1e20, lo > -1e20
    Aub = np.vstack([np.c_[A[fu], -np.ones(fu.sum())], np.c_[-A[fl], -np.ones(fl.sum())]])
    r = linprog(np.r_[np.zeros(n), 1.0], A_ub=Aub, b_ub=np.r_[hi[fu], -lo[fl]], bounds=[(None, None)] * n + [(0, None)], method="highs")
    assert r.status == 0 and r.fun > 1e-5   # smallest uniform violation of l <= A z <= u is positive: infeasible
    # the whole SQP cycle on it: the QP fails, the step stays zero, the reference's loop then reports SOLVED (quirk 10)
    r2 = emu.warp_solve_ocp(pf, table, p["Ts"], N, g["warm"], g["rb"], g["u"])
    assert r2["status"] == 0 and r2["iters"] == 1 and r2["qp_ok"][0] == 0 and r2["qp_iters"] < 30


# ---- second-order correction (sqp.json "do_SOC", osqp_interface.cpp:506-533,658-681) ----------------------------
@pytest.mark.parametrize("N,lanes", [(10, 32), (20, 32), (10, 128)])
def test_second_order_correction_vs_oracle(O, nn, emu, track_wp, rng, N, lanes):
    """solveOCP with do_SOC: after every QP a second QP with the same P, q, A and bounds shifted by d = c(x (+) step) - A step replaces the step.
    The oracle restates that with the reference's dense matrices (setConstraints at the shifted point, jac * step); the product derives the
    shifted bounds per stage in closed form (sqp_warp.cuh::soc_right_hand_sides).  Same statuses, iteration counts and line-search decisions
    (replayed along the product's branch, ties certified), steps within the QP tolerance -- and the correction must actually change the steps
    (compared with the same solve without it), otherwise the test would pass on a no-op."""
    p = O.load_params(overrides={"sqp": {"do_SOC": True}})
    assert p["sqp"][4] == 1.0
    pf = flat_params(p)
    p0 = O.load_params(); pf0 = flat_params(p0)
    table = emu.fit_track(*track_wp)
    Ts = p["Ts"]
    o = O.OracleMPC(N=N, nn=nn, params=p)
    o.set_track(*track_wp)
    changed = 0.0
    for trial in range(4):
        q0 = O.Q_HOME + rng.uniform(-0.05, 0.05, 7)
        hor = np.tile(np.r_[q0, 0.0, 0.0, np.zeros(8)], (N + 1, 1))
        if trial >= 2:   # a warm-start-like guess with non-zero inputs and a moving path parameter
            hor[:, 7] = np.linspace(0.0, 0.02, N + 1); hor[:, 8] = 0.1; hor[:N, 9:16] = rng.uniform(-0.05, 0.05, (N, 7)); hor[:N, 16] = 0.2
        rb = np.stack([nn.robot_data(hor[k, :7]) for k in range(N + 1)])
        cur_u = np.zeros(8) if trial < 2 else np.r_[hor[0, 9:16], 0.0]
        a = emu.warp_solve_ocp(pf, table, Ts, N, hor, rb, cur_u, reverse=bool(trial % 2), lanes=lanes, soc=True)
        o.set_forced_decisions([int(x == 1.0) for x in a["alphas"]])
        b = o.solve_ocp(hor, rb, cur_u)
        nat, mg = o.decision_log()
        o.set_forced_decisions([])
        assert a["status"] == b["status"] == 0 and a["iters"] == b["iters"], (a["status"], b["status"], a["iters"], b["iters"])
        assert np.allclose(a["alphas"], b["alphas"])
        for i in range(len(b["steps"])):
            assert np.abs(step_to_flat(a["steps"][i], N) - b["steps"][i]).max() < 1e-4, (trial, i, np.abs(step_to_flat(a["steps"][i], N) - b["steps"][i]).max())
            assert nat[i] == int(a["alphas"][i] == 1.0) or mg[i] < 1e-6
        assert np.abs(a["horizon"] - b["horizon"]).max() < 1e-4
        # the instantiation with the correction compiled in but switched off is the plain loop
        c = emu.warp_solve_ocp(pf0, table, Ts, N, hor, rb, cur_u, lanes=lanes, soc=True)
        d = emu.warp_solve_ocp(pf0, table, Ts, N, hor, rb, cur_u, lanes=lanes)
        assert c["iters"] == d["iters"] and np.array_equal(c["horizon"], d["horizon"])
        changed = max(changed, np.abs(a["steps"][0] - d["steps"][0]).max())
    assert changed > 1e-3, changed
