"""The reference's Python surface (python/MPCC/MPCC.py, robot_model.py over MPCC_WRAPPER) on the CUDA path, and the live
setParam semantics."""
import numpy as np
import pytest

TU = np.array([2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61, 5.0])


def test_live_override_rule():
    """CPU: which maps of a ParamValue take effect on a live controller (osqp_interface.cpp:95-100)."""
    from mpcc_manipulator_b200.MPCC import live_overrides, _check_param_value
    pv = {"param": {"tol_sing": 0.018}, "cost": {"qC": 800.0}, "bounds": {"q1l": -1.0}, "normalization": {"q1": 2.0}, "sqp": {"eps_prim": 1e-9}}
    _check_param_value(pv)
    assert live_overrides(pv) == {"model.tol_sing": 0.018, "cost.qC": 800.0}
    with pytest.raises(AssertionError):
        _check_param_value({"nofile": {}})
    with pytest.raises(AssertionError):
        _check_param_value({"cost": {"nokey": 1.0}})


def test_oracle_live_set_param_semantics(O, nn, track_wp):
    """CPU: the oracle's restatement of MPC::setParam: cost / model maps act, bounds / normalisation / sqp do not."""
    a = O.OracleMPC(N=10, nn=nn); a.set_track(*track_wp)
    b = O.OracleMPC(N=10, nn=nn, params=O.load_params(overrides={"cost": {"qC": 800.0}, "model": {"desired_ee_velocity": 0.1}})); b.set_track(*track_wp)
    a.set_param_live({"cost": {"qC": 800.0}, "param": {"desired_ee_velocity": 0.1}, "sqp": {"eps_prim": 1e-9, "max_iter": 2}, "bounds": {"q1u": 0.0}, "normalization": {"q1": 9.0}})
    x = np.r_[O.Q_HOME, 0., 0.]; u = np.zeros(8)
    ra, rb = a.run(x, u), b.run(x, u)
    assert ra["status"] == rb["status"] == 0 and ra["iters"] == rb["iters"] and np.array_equal(ra["u0"], rb["u0"])


@pytest.mark.gpu
def test_mpcc_class_matches_reference_surface(O, nn, track_wp):
    """python/main.py-style use: MPCC(), setTrack(state), runMPC(state, input) -> (status, state, input, horizon, times)."""
    from mpcc_manipulator_b200.MPCC import MPCC
    mpcc = MPCC()
    assert mpcc.pred_horizon == 10 and mpcc.robot_dof == 7 and mpcc.num_links == 9 and mpcc.Ts == 0.01
    state = np.r_[O.Q_HOME, 0., 0.]
    ee = mpcc.robot_model.getEEPosition(O.Q_HOME)
    assert np.abs(ee - np.array([0.5545, 0.0, 0.5211])).max() < 1e-4      # python/main_utils.py:50-52
    J = mpcc.robot_model.getEEJacobian(O.Q_HOME)
    p, R, Jo = O.fk(O.Q_HOME)
    assert J.shape == (6, 7) and np.abs(J - Jo).max() < 1e-12 and np.abs(mpcc.robot_model.getEEOrientation(O.Q_HOME) - R).max() < 1e-12
    assert abs(mpcc.robot_model.getEEManipulability(O.Q_HOME) - O.manip(O.Q_HOME)[0]) < 1e-12
    mpcc.setTrack(state)
    pos, rot, arc = mpcc.getSplinePath()
    assert pos.shape == (100, 3) and rot.shape == (100, 3, 3) and arc.shape == (100,) and np.abs(pos[0] - ee).max() < 1e-12
    o = O.OracleMPC(N=10, nn=nn); o.set_track(*track_wp)
    rp, rR = mpcc.getRefPose(0.3)
    te = o.track_eval(0.3)
    assert np.abs(rp - te["pos"]).max() < 1e-12 and np.abs(rR - te["R"]).max() < 1e-12
    assert abs(mpcc.getContourError(0.0, ee)) < 1e-12
    x, u = state.copy(), np.zeros(8)
    for c in range(3):
        w_hor, w_valid, w_failed = mpcc.mpc.get_warm_state()
        ok, x_upd, u_new, hor, ct = mpcc.runMPC(x, u)
        r = mpcc.mpc.read_results()
        dec = [(int(mpcc.mpc.decisions()[0]) >> i) & 1 for i in range(min(int(r["iters"][0]), 32))]
        o.set_warm_state(w_hor[0], w_valid[0], w_failed[0]); o.set_forced_decisions(dec)      # along the device's line-search branch (filter ties)
        ro = o.run(x, u)
        assert ok is True and ro["ok"] and len(hor) == 11 and set(hor[0]) == {"state", "input"} and set(ct) == {"total", "set_qp", "solve_qp", "get_alpha", "set_env"}
        assert np.abs(x_upd - ro["x0"]).max() < 1e-9 and ct["total"] > 0
        assert r["iters"][0] == ro["iters"] and (np.abs(u_new - ro["u0"]) / TU).max() < 1e-4
        assert np.array_equal(hor[0]["input"], u_new)
        u = u_new; x = O.sim_time_step(x_upd, u_new, 0.01)
    mpcc.close()


@pytest.mark.gpu
def test_live_set_param_mid_run(O, nn, track_wp):
    """main.cpp:103-106: setParam in the middle of a run.  Warm starts survive; the cost / model maps act from the next cycle
    on; sqp / bounds / normalisation maps are ignored (quirk 13).  GPU against the oracle's restatement, cycle by cycle."""
    from mpcc_manipulator_b200.MPCC import MPCC
    pv = {"cost": {"qC": 800.0, "qVs": 10.0}, "param": {"desired_ee_velocity": 0.1}, "sqp": {"max_iter": 1, "eps_prim": 1e-12}, "bounds": {"q1u": 0.0}}
    mpcc = MPCC()
    o = O.OracleMPC(N=10, nn=nn); o.set_track(*track_wp)
    x = np.r_[O.Q_HOME, 0., 0.]; u = np.zeros(8)
    mpcc.setTrack(x)
    n_cmp = 0
    for c in range(8):
        if c == 4:
            mpcc.setParam(pv); o.set_param_live(pv)
        w_hor, w_valid, w_failed = mpcc.mpc.get_warm_state()
        if c == 4:
            assert w_valid[0] == 1                                        # the warm start is kept across setParam
        ok, x_upd, u_new, hor, ct = mpcc.runMPC(x, u)
        r = mpcc.mpc.read_results()
        dec = [(int(mpcc.mpc.decisions()[0]) >> i) & 1 for i in range(min(int(r["iters"][0]), 32))]
        o.set_warm_state(w_hor[0], w_valid[0], w_failed[0]); o.set_forced_decisions(dec)
        ro = o.run(x, u)
        assert ok and r["status"][0] == ro["status"] == 0 and r["iters"][0] == ro["iters"]
        assert (np.abs(u_new - ro["u0"]) / TU).max() < 1e-4
        n_cmp += 1
        u = u_new; x = O.sim_time_step(x_upd, u_new, 0.01)
    assert n_cmp == 8
    mpcc.close()
