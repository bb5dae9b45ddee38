"""CPU tier: the committed golden fixtures (tests/golden, made by tests/gen_golden.py from the oracle) still match
the oracle bit-for-bit-ish (regression pin of the checker itself) and the product's host-compiled device code."""
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from helpers import Emul, flat_params, assemble_flat_qp_from_lin, step_to_flat

G = Path(__file__).resolve().parent / "golden"


def test_oracle_reproduces_robot_data(O, nn):
    g = np.load(G / "robot_data.npz")
    for i in range(0, len(g["q"]), 7):
        rb = nn.robot_data(g["q"][i], g["obs"][i])
        assert np.abs(rb - g["rb"][i]).max() <= 1e-12 * np.abs(g["rb"][i]).max()


def test_oracle_reproduces_closed_loop(O, nn, track_wp):
    g = np.load(G / "closed_loop_c1.npz")
    o = O.OracleMPC(N=10, nn=nn); o.set_track(*track_wp)
    for c in range(6):
        r = o.run(g["x_in"][c], g["u_in"][c])
        assert r["status"] == g["status"][c] and r["iters"] == g["iters"][c]
        assert np.abs(r["u0"] - g["u_out"][c]).max() < 1e-9
        assert np.abs(r["x0"] - g["x_out"][c]).max() < 1e-12


def test_product_host_code_vs_golden_flat_qp(O):
    g = np.load(G / "flat_qp_n10.npz")
    emu = Emul()
    ee = O.fk(O.Q_HOME)[0]
    X, Y, Z, R = O.load_track(); X, Y, Z = O.shift_track(X, Y, Z, ee)
    table = emu.fit_track(X, Y, Z, R)
    p = O.load_params(); pf = flat_params(p); Ts, N = p["Ts"], 10
    hor, rb, cur_u = g["hor"], g["rb"], g["cur_u"]
    lin = []
    for k in range(N + 1):
        up = cur_u[:7] if k == 0 else hor[k - 1, 9:16]
        un = hor[k + 1, 9:16] if k < N else np.zeros(7)
        xn = hor[k + 1, :9] if k < N else np.zeros(9)
        lin.append(emu.stage_lin(pf, table, Ts, N, k, hor[k, :9], hor[k, 9:], up, un, xn, rb[k]))
    P, q = assemble_flat_qp_from_lin(np.array(lin), pf, N, Ts)
    assert np.abs(P - g["P"]).max() <= 1e-9 * np.abs(g["P"]).max()
    assert np.abs(q - g["q"]).max() <= 1e-9 * np.abs(g["q"]).max()
    ok, step, it, res = emu.solve_qp(pf, table, Ts, N, hor, rb, cur_u)
    assert ok and np.abs(step_to_flat(step, N) - g["z"]).max() < 1e-4
