"""Shared test helpers (test infrastructure; may use the oracle)."""
import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
NX, NU, HZ, RB, LIN, DOF, NPC = 9, 8, 17, 150, 212, 7, 11
# StageLin layout (csrc/dev_qp.cuh)
LQ, Lq, LRD, Lr, Lb, LXLO, LXHI, LDLO, LDHI, LPG, LPD, LPRHS, LOBJ, LGAP = 0, 45, 54, 62, 70, 79, 88, 97, 104, 111, 188, 199, 210, 211


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emul:
    """The product's device functions compiled for the host (tests/emul/libemul.so)."""

    def __init__(self):
        self.lib = C.CDLL(str(ROOT / "tests" / "emul" / "libemul.so"))
        self.lib.emu_project.restype = C.c_double

    def fit_track(self, X, Y, Z, R):
        t = np.zeros(self.lib.emu_track_doubles())
        X, Y, Z, R = f64(X), f64(Y), f64(Z), f64(R)
        self.lib.emu_fit_track(len(X), _p(X), _p(Y), _p(Z), _p(R), _p(t))
        return t

    def kin(self, q):
        o = np.zeros(62)
        self.lib.emu_kin(_p(f64(q)), _p(o))
        return dict(p=o[0:3], R=o[3:12].reshape(3, 3), Jv=o[12:33].reshape(3, 7), Jw=o[33:54].reshape(3, 7), manip=o[54], dmanip=o[55:62])

    def track_eval(self, table, s):
        o = np.zeros(21)
        self.lib.emu_track_eval(_p(table), C.c_double(s), _p(o))
        return o

    def project(self, table, max_dist, s, ee):
        return self.lib.emu_project(_p(table), C.c_double(max_dist), C.c_double(s), _p(f64(ee)))

    def stage_lin(self, params, table, Ts, N, k, x, u, up, un, xn, rb):
        o = np.zeros(LIN)
        self.lib.emu_stage_lin(_p(params), _p(table), C.c_double(Ts), N, k, _p(f64(x)), _p(f64(u)), _p(f64(up)), _p(f64(un)), _p(f64(xn)), _p(f64(rb)), _p(o))
        return o

    def so3(self, R):
        lg, ex = np.zeros(3), np.zeros(9)
        self.lib.emu_so3(_p(f64(R)), _p(lg), _p(ex))
        return lg, ex.reshape(3, 3)

    def prologue(self, params, table, Ts, N, x0, u0, warm, valid, failed):
        x0 = f64(x0).copy(); warm = f64(warm).copy()
        v, f = C.c_int(valid), C.c_int(failed)
        self.lib.emu_prologue(_p(params), _p(table), C.c_double(Ts), N, _p(x0), _p(f64(u0)), _p(warm), C.byref(v), C.byref(f))
        return x0, warm, v.value, f.value

    def solve_ocp(self, params, table, Ts, N, guess, rb, cur_u, qp_max_iter=60, qp_eps=1e-9, max_log=100):
        g = f64(guess).copy()
        st, it, qi, nl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        steps = np.zeros((max_log, N + 1, HZ)); al = np.zeros(max_log); ok_ = np.zeros(max_log, np.int32)
        ok = self.lib.emu_solve_ocp(_p(params), _p(table), C.c_double(Ts), N, _p(g), _p(f64(rb)), _p(f64(cur_u)), qp_max_iter, C.c_double(qp_eps),
                                    C.byref(st), C.byref(it), C.byref(qi), _p(steps), _p(al), _p(ok_), max_log, C.byref(nl))
        k = nl.value
        return dict(ok=bool(ok), horizon=g, status=st.value, iters=it.value, qp_iters=qi.value, steps=steps[:k], alphas=al[:k], qp_ok=ok_[:k])

    def epilogue(self, N, status, iters, x0, guess, valid, failed):
        g = f64(guess).copy(); v, f = C.c_int(valid), C.c_int(failed)
        ok = self.lib.emu_epilogue(N, int(status), int(iters), _p(f64(x0)), _p(g), C.byref(v), C.byref(f))
        return bool(ok), g, v.value, f.value

    def run_cycle(self, nn, params, table, Ts, N, x0, u0, state, obs=(3., 3., 3., 0.), qp_max_iter=60, qp_eps=1e-9):
        """One whole control cycle with the product's host-compiled code; RobotData from the oracle's networks
        (the GPU tier checks the CUDA RobotData against the same oracle at 1e-9).  state = [warm, valid, failed]."""
        x0n, warm, v, f = self.prologue(params, table, Ts, N, x0, u0, state[0], state[1], state[2])
        rb = np.stack([nn.robot_data(warm[k, :7], obs) for k in range(N + 1)])
        r = self.solve_ocp(params, table, Ts, N, warm, rb, u0, qp_max_iter, qp_eps)
        ok, g, v, f = self.epilogue(N, r["status"], r["iters"], x0n, r["horizon"], v, f)
        state[0], state[1], state[2] = g, v, f
        return dict(x0=x0n, u0=g[0, 9:].copy(), horizon=g, status=r["status"], iters=r["iters"], ok=ok, accept=[int(a == 1.0) for a in r["alphas"]], qp_ok=r["qp_ok"])

    def warp_solve_ocp(self, params, table, Ts, N, guess, rb, cur_u, qp_max_iter=60, qp_eps=1e-9, max_log=100, reverse=False, lanes=32, soc=False):
        """lanes = 32: the warp kernel's code; lanes = 128: the CTA (latency-mode) kernel's code; soc: the instantiation with the second-order
        correction compiled in (k_sqp_soc.cu's kernels; sqp.do_SOC in `params` switches it on)"""
        g = f64(guess).copy()
        st, it, qi, nl, am = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_uint()
        steps = np.zeros((max_log, N + 1, HZ)); al = np.zeros(max_log); ok_ = np.zeros(max_log, np.int32)
        fn = {(32, False): self.lib.emu_warp_solve_ocp, (128, False): self.lib.emu_cta_solve_ocp, (32, True): self.lib.emu_warp_solve_ocp_soc, (128, True): self.lib.emu_cta_solve_ocp_soc}[(lanes, bool(soc))]
        ok = fn(_p(params), _p(table), C.c_double(Ts), N, _p(g), _p(f64(rb)), _p(f64(cur_u)), qp_max_iter, C.c_double(qp_eps), int(reverse),
                                         C.byref(st), C.byref(it), C.byref(qi), _p(steps), _p(al), _p(ok_), max_log, C.byref(nl), C.byref(am))
        k = nl.value
        return dict(ok=bool(ok), horizon=g, status=st.value, iters=it.value, qp_iters=qi.value, steps=steps[:k], alphas=al[:k], qp_ok=ok_[:k], accept_mask=am.value)

    def warp_solve_qp(self, params, table, Ts, N, guess, rb, cur_u, qp_max_iter=60, qp_eps=1e-9, reverse=False):
        step = np.zeros((N + 1, HZ)); it = C.c_int(); res = np.zeros(3)
        ok = self.lib.emu_warp_solve_qp(_p(params), _p(table), C.c_double(Ts), N, _p(f64(guess)), _p(f64(rb)), _p(f64(cur_u)), qp_max_iter, C.c_double(qp_eps),
                                        int(reverse), _p(step), C.byref(it), _p(res))
        return bool(ok), step, it.value, res

    def solve_qp(self, params, table, Ts, N, guess, rb, cur_u, qp_max_iter=60, qp_eps=1e-9):
        step = np.zeros((N + 1, HZ)); it = C.c_int(); res = np.zeros(3)
        ok = self.lib.emu_solve_qp(_p(params), _p(table), C.c_double(Ts), N, _p(f64(guess)), _p(f64(rb)), _p(f64(cur_u)), qp_max_iter, C.c_double(qp_eps),
                                   _p(step), C.byref(it), _p(res))
        return bool(ok), step, it.value, res


def flat_params(p):
    """oracle.load_params() dict -> the product's flat params record (include/mpcc_cuda.h order)."""
    return f64(np.concatenate([p["model"], p["cost"], p["bounds"], p["norm"], p["sqp"], [p["cost"][7]]]))


def step_to_flat(step_hz, N):
    """horizon-layout step [N+1][17] -> the reference's flat z (all states, then all inputs; osqp_interface.cpp:117-118)."""
    s = np.asarray(step_hz)
    return np.concatenate([s[:, :NX].ravel(), s[:N, NX:].ravel()])


def unpack_sym9(v45):
    Q = np.zeros((9, 9)); q = 0
    for r in range(9):
        for c in range(r + 1):
            Q[r, c] = Q[c, r] = v45[q]; q += 1
    return Q


def assemble_flat_qp_from_lin(lin, params_flat, N, Ts):
    """Re-assemble the reference's flat QP objective from the product's per-stage blocks.
    Returns P (n_var x n_var), q (n_var).  Used to compare with the oracle's osqp_interface-style dense assembly."""
    Tu = params_flat[7 + 12 + 48 + 9: 7 + 12 + 48 + 17]
    r_ddq = params_flat[-1]
    nx = NX * (N + 1); n = nx + NU * N
    P = np.zeros((n, n)); q = np.zeros(n)
    for k in range(N + 1):
        L = lin[k]
        P[9 * k:9 * k + 9, 9 * k:9 * k + 9] = unpack_sym9(L[LQ:LQ + 45])
        q[9 * k:9 * k + 9] = L[Lq:Lq + 9]
        if k < N:
            o = nx + 8 * k
            P[o:o + 8, o:o + 8] = np.diag(L[LRD:LRD + 8])
            q[o:o + 8] = L[Lr:Lr + 8]
            if k + 1 < N:
                for j in range(DOF):
                    P[o + j, o + 8 + j] = P[o + 8 + j, o + j] = -2.0 * r_ddq * Tu[j] * Tu[j]
    return P, q


def make_horizon(O, rng, N, Ts, q0=None, spread=0.05, u_scale=0.2):
    """A dynamically consistent random horizon around q_home: [N+1][17]."""
    q = (O.Q_HOME if q0 is None else q0) + rng.uniform(-spread, spread, 7)
    x = np.r_[q, 0.02, 0.1]
    hor = np.zeros((N + 1, HZ))
    for k in range(N + 1):
        hor[k, :NX] = x
        if k < N:
            u = np.r_[rng.uniform(-u_scale, u_scale, 7), rng.uniform(-0.5, 0.5)]
            hor[k, NX:] = u
            x = O.rk4(x, u, Ts)
    return hor
