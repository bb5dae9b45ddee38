"""CPU tier: whole control cycles (prologue -> SQP -> epilogue) of the product's host-compiled code, closed loop,
with the oracle replayed along the product's line-search branch (certified filter ties)."""
import numpy as np
import pytest

from helpers import Emul, flat_params

TIE, QP_TOL = 1e-6, 1e-4


@pytest.mark.parametrize("N,cycles,perturb", [(10, 15, False), (20, 4, True)])
def test_closed_loop_follow(O, nn, track_wp, N, cycles, perturb):
    emu = Emul()
    p = O.load_params(); pf = flat_params(p); Ts = p["Ts"]
    Tu = p["norm"][9:]
    table = emu.fit_track(*track_wp)
    o = O.OracleMPC(N=N, nn=nn); o.set_track(*track_wp)
    rng = np.random.default_rng(0)
    x = np.r_[O.Q_HOME + (rng.uniform(-0.05, 0.05, 7) if perturb else 0), 0., 0.]; u = np.zeros(8)
    state = [np.zeros((N + 1, 17)), 0, 0]
    ties = 0
    for c in range(cycles):
        r = emu.run_cycle(nn, pf, table, Ts, N, x, u, state)
        o.set_forced_decisions(r["accept"])
        ro = o.run(x, u)
        nat, mg = o.decision_log()
        assert r["status"] == ro["status"] == 0 and r["iters"] == ro["iters"] and r["ok"] == ro["ok"]
        assert np.abs(r["x0"] - ro["x0"]).max() < 1e-12
        assert (np.abs(r["u0"] - ro["u0"]) / Tu).max() < QP_TOL
        for i in range(len(r["accept"])):
            if nat[i] != r["accept"][i]:
                assert mg[i] < TIE
                ties += 1
        hor, valid, failed = o.warm_state()
        assert valid == state[1] and failed == state[2]
        u = r["u0"]; x = O.sim_time_step(r["x0"], u, Ts)
    assert x[7] > 0.0  # progress along the path
