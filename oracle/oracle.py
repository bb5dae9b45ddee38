"""TEST INFRASTRUCTURE ONLY: ctypes front-end of the CPU oracle (oracle/liborc.so).

Only tests/, bench.py's cpu_baseline / --impl reference leg and __graft_entry__.smoke()
may import this module, and only as the checker.  The product package never does.
"""
import ctypes as C
import json
import struct
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ASSETS = HERE.parent / "mpcc_manipulator_b200" / "assets"
NX, NU, NPC, DOF, NLINKS, RB = 9, 8, 11, 7, 9, 150

MODEL_KEYS = ["max_dist_proj", "desired_ee_velocity", "s_trust_region", "deaccelerate_ratio", "tol_sing", "tol_selcol", "tol_envcol"]
COST_KEYS = ["qC", "qCNmult", "qL", "qVs", "qOri", "qSing", "rdq", "rddq", "rdVs", "qC_reduction_ratio", "qL_increase_ratio", "qOri_reduction_ratio"]
X_NAMES = ["q1", "q2", "q3", "q4", "q5", "q6", "q7", "s", "vs"]
U_NAMES = ["dq1", "dq2", "dq3", "dq4", "dq5", "dq6", "dq7", "dVs"]
DD_NAMES = ["ddq%d" % i for i in range(1, 8)]
SQP_KEYS = ["eps_prim", "eps_dual", "max_iter", "line_search_max_iter", "do_SOC", "use_BFGS", "line_search_tau", "line_search_eta", "line_search_rho"]


def build():
    subprocess.check_call(["make", "-s", "-C", str(HERE)])


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = HERE / "liborc.so"
        if not so.exists():
            build()
        _lib = C.CDLL(str(so))
        _lib.orc_nn_create.restype = C.c_void_p
        _lib.orc_mpc_create.restype = C.c_void_p
        _lib.orc_mpc_track_length.restype = C.c_double
        _lib.orc_project.restype = C.c_double
        _lib.orc_rbf.restype = C.c_double
        _lib.orc_mpc_last_filter_margin.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def load_params(param_dir=None, overrides=None):
    """Reference JSON schema (cpp/Params/*.json, params.cpp) -> flat arrays in the oracle's order."""
    d = Path(param_dir) if param_dir else ASSETS / "params"
    ov = overrides or {}

    def jl(name):
        v = json.loads((d / f"{name}.json").read_text())
        v.update(ov.get(name, {}))
        return v

    m, c, b, n, s = jl("model"), jl("cost"), jl("bounds"), jl("normalization"), jl("sqp")
    cfg = json.loads((d / "config.json").read_text())
    out = {
        "Ts": float(cfg["Ts"]),
        "model": f64([m[k] for k in MODEL_KEYS]),
        "cost": f64([c[k] for k in COST_KEYS]),
        "bounds": f64([b[k + "l"] for k in X_NAMES] + [b[k + "u"] for k in X_NAMES] + [b[k + "l"] for k in U_NAMES]
                      + [b[k + "u"] for k in U_NAMES] + [b[k + "l"] for k in DD_NAMES] + [b[k + "u"] for k in DD_NAMES]),
        "norm": f64([n[k] for k in X_NAMES] + [n[k] for k in U_NAMES]),
        "sqp": f64([float(s[k]) for k in SQP_KEYS]),
    }
    return out


def load_track(path=None):
    """track.json -> X, Y, Z, R (n x 9 row-major) as Track::Track does (track.cpp:19-54)."""
    t = json.loads(Path(path or ASSETS / "params" / "track.json").read_text())
    X, Y, Z = f64(t["X"]), f64(t["Y"]), f64(t["Z"])
    q = np.stack([f64(t["quat_X"]), f64(t["quat_Y"]), f64(t["quat_Z"]), f64(t["quat_W"])], 1)
    return X, Y, Z, quat_to_rot(q)


def quat_to_rot(q):
    """Eigen Quaterniond(x,y,z,w).normalized().toRotationMatrix(), rows flattened."""
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.stack([1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)], 1)
    return f64(R)


def shift_track(X, Y, Z, init_position):
    """Track::getTrack (track.cpp:56-66): translate so the first point is init_position."""
    return X - X[0] + init_position[0], Y - Y[0] + init_position[1], Z - Z[0] + init_position[2]


def read_nn(path):
    blob = Path(path).read_bytes()
    assert blob[:8] == b"MPCCNN1\0"
    (nl,) = struct.unpack_from("<i", blob, 8)
    dims = [struct.unpack_from("<ii", blob, 12 + 8 * k) for k in range(nl)]
    off = 12 + 8 * nl
    Ws, bs = [], []
    for o, i in dims:
        Ws.append(np.frombuffer(blob, "<f8", o * i, off).reshape(o, i).copy()); off += 8 * o * i
        bs.append(np.frombuffer(blob, "<f8", o, off).copy()); off += 8 * o
    return Ws, bs


class OracleNN:
    def __init__(self, nn_dir=None):
        d = Path(nn_dir) if nn_dir else ASSETS / "nn"
        self.self_W, self.self_b = read_nn(d / "self_collision.f64")
        self.env_W, self.env_b = read_nn(d / "env_collision.f64")
        cat = lambda xs: f64(np.concatenate([x.ravel() for x in xs]))
        self._keep = [cat(self.self_W), cat(self.self_b), cat(self.env_W), cat(self.env_b)]
        self.h = C.c_void_p(lib().orc_nn_create(*[_p(a) for a in self._keep]))

    def mlp(self, which, x):
        x = f64(x)
        no, ni = (1, 7) if which == 0 else (9, 10)
        out, jac = np.zeros(no), np.zeros((no, ni))
        lib().orc_mlp_eval(self.h, which, _p(x), _p(out), _p(jac))
        return out, jac

    def robot_data(self, q, obs4=(3., 3., 3., 0.)):
        rb = np.zeros(RB)
        lib().orc_robot_data(self.h, _p(f64(q)), _p(f64(obs4)), _p(rb))
        return rb


class OracleMPC:
    """One reference-style MPC object (mpc.h:58-101) on the CPU."""

    def __init__(self, N=10, params=None, nn=None, Ts=None):
        self.N = N
        self.params = params or load_params()
        self.nn = nn or OracleNN()
        self.Ts = Ts if Ts is not None else self.params["Ts"]
        p = self.params
        self.h = C.c_void_p(lib().orc_mpc_create(N, C.c_double(self.Ts), self.nn.h, _p(p["model"]), _p(p["cost"]), _p(p["bounds"]), _p(p["norm"]), _p(p["sqp"])))
        self.n_var = (N + 1) * NX + N * NU
        self.n_constr = (N + 1) * NX + (self.n_var + N * NU) + (N + 1) * NPC

    def __del__(self):
        try:
            lib().orc_mpc_destroy(self.h)
        except Exception:
            pass

    def set_params(self, params):
        self.params = params
        p = params
        lib().orc_mpc_set_params(self.h, _p(p["model"]), _p(p["cost"]), _p(p["bounds"]), _p(p["norm"]), _p(p["sqp"]))

    def set_param_live(self, param_value, param_dir=None):
        """MPC::setParam(ParamValue) on a live controller with the reference's semantics (quirk 13): only the "param" (model)
        and "cost" maps take effect; bounds come from the FILE; normalisation, SQP parameters and the interface's rddq stay."""
        p = load_params(param_dir, overrides={"model": dict(param_value.get("param", {})), "cost": dict(param_value.get("cost", {}))})
        lib().orc_mpc_set_param_live(self.h, _p(p["model"]), _p(p["cost"]), _p(p["bounds"]))

    def set_qp_options(self, max_iter=100, eps=1e-9):
        lib().orc_mpc_set_qp_options(self.h, max_iter, C.c_double(eps))

    def set_track(self, X, Y, Z, R):
        X, Y, Z, R = f64(X), f64(Y), f64(Z), f64(R)
        lib().orc_mpc_set_track(self.h, len(X), _p(X), _p(Y), _p(Z), _p(R))

    def reset(self):
        lib().orc_mpc_reset(self.h)

    @property
    def track_length(self):
        return lib().orc_mpc_track_length(self.h)

    def run(self, x0, u0, obs4=(3., 3., 3., 0.)):
        x0 = f64(x0).copy()
        u = np.zeros(NU); hor = np.zeros((self.N + 1, 17)); st = C.c_int(); it = C.c_int(); tm = np.zeros(5)
        ok = lib().orc_mpc_run(self.h, _p(x0), _p(f64(u0)), _p(f64(obs4)), _p(u), _p(hor), C.byref(st), C.byref(it), _p(tm))
        return dict(ok=bool(ok), x0=x0, u0=u, horizon=hor, status=st.value, iters=it.value, times=tm)

    def last_filter_margin(self):
        """Smallest robustness margin of the filter accept/reject comparisons in the last solveOCP (1e300: none made)."""
        return lib().orc_mpc_last_filter_margin(self.h)

    def set_forced_decisions(self, accept):
        """Follow these per-iteration first-trial accept (1) / reject (0) decisions instead of the filter's own."""
        a = np.ascontiguousarray(accept, dtype=np.int32)
        lib().orc_mpc_set_forced_decisions(self.h, _p(a), len(a))

    def decision_log(self, max_n=128):
        nat = np.zeros(max_n, np.int32); mg = np.zeros(max_n)
        n = lib().orc_mpc_decision_log(self.h, _p(nat), _p(mg), max_n)
        return nat[:n], mg[:n]

    def warm_state(self):
        hor = np.zeros((self.N + 1, 17)); v = C.c_int(); f = C.c_int()
        lib().orc_mpc_warm_state(self.h, _p(hor), C.byref(v), C.byref(f))
        return hor, v.value, f.value

    def set_qp_eps(self, eps):
        """Termination threshold of the checker's dense QP solver (all KKT residuals; default 1e-9)."""
        lib().orc_mpc_set_qp_eps(self.h, C.c_double(eps))

    def set_warm_state(self, horizon, valid, failed):
        """Replay harness: start the next cycle from this warm start (e.g. the one the CUDA path holds)."""
        hor = np.ascontiguousarray(horizon, dtype=np.float64)
        assert hor.shape == (self.N + 1, 17)
        lib().orc_mpc_set_warm_state(self.h, _p(hor), int(valid), int(failed))

    def track_eval(self, s):
        o = np.zeros(21)
        lib().orc_track_eval(self.h, C.c_double(s), _p(o))
        return dict(pos=o[0:3], dpos=o[3:6], ddpos=o[6:9], R=o[9:18].reshape(3, 3), dR=o[18:21])

    def track_table(self):
        s, X, Y, Z, R = np.zeros(100), np.zeros(100), np.zeros(100), np.zeros(100), np.zeros((100, 9))
        lib().orc_track_table(self.h, _p(s), _p(X), _p(Y), _p(Z), _p(R))
        return s, X, Y, Z, R

    def project(self, s, ee):
        return lib().orc_project(self.h, C.c_double(s), _p(f64(ee)))

    def stage_cost(self, x, u, rb, k):
        obj = C.c_double(); fx, fu, fxx, fuu = np.zeros(9), np.zeros(8), np.zeros((9, 9)), np.zeros((8, 8))
        lib().orc_stage_cost(self.h, _p(f64(x)), _p(f64(u)), _p(f64(rb)), k, C.byref(obj), _p(fx), _p(fu), _p(fxx), _p(fuu))
        return obj.value, fx, fu, fxx, fuu

    def stage_constraints(self, x, u, rb, k):
        c, cl, cu, cx, cuu = np.zeros(11), np.zeros(11), np.zeros(11), np.zeros((11, 9)), np.zeros((11, 8))
        lib().orc_stage_constraints(self.h, _p(f64(x)), _p(f64(u)), _p(f64(rb)), k, _p(c), _p(cl), _p(cu), _p(cx), _p(cuu))
        return c, cl, cu, cx, cuu

    def build_qp(self, guess, rb, cur_u):
        n, m = self.n_var, self.n_constr
        P, q, A = np.zeros((n, n)), np.zeros(n), np.zeros((m, n))
        l, u, c = np.zeros(m), np.zeros(m), np.zeros(m); obj = C.c_double()
        lib().orc_build_qp(self.h, _p(f64(guess)), _p(f64(rb)), _p(f64(cur_u)), _p(P), _p(q), _p(A), _p(l), _p(u), _p(c), C.byref(obj))
        return dict(P=P, q=q, A=A, l=l, u=u, c=c, obj=obj.value)

    def solve_ocp(self, guess, rb, cur_u, max_log=100):
        g = f64(guess).copy(); st = C.c_int(); it = C.c_int(); nl = C.c_int()
        steps = np.zeros((max_log, self.n_var)); alphas = np.zeros(max_log); ok_ = np.zeros(max_log, dtype=np.int32)
        ok = lib().orc_solve_ocp(self.h, _p(g), _p(f64(rb)), _p(f64(cur_u)), C.byref(st), C.byref(it), _p(steps), _p(alphas), _p(ok_), max_log, C.byref(nl))
        k = nl.value
        return dict(ok=bool(ok), horizon=g, status=st.value, iters=it.value, steps=steps[:k], alphas=alphas[:k], qp_ok=ok_[:k])


def solve_qp_dense(P, q, A, l, u, max_iter=100, eps=1e-9):
    n, m = len(q), len(l)
    z = np.zeros(n); it = C.c_int()
    ok = lib().orc_solve_qp_dense(n, m, _p(f64(P)), _p(f64(q)), _p(f64(A)), _p(f64(l)), _p(f64(u)), _p(z), C.byref(it), max_iter, C.c_double(eps))
    return bool(ok), z, it.value


def fk(q):
    p, R, J = np.zeros(3), np.zeros((3, 3)), np.zeros((6, 7))
    lib().orc_fk(_p(f64(q)), _p(p), _p(R), _p(J))
    return p, R, J


def manip(q):
    m = C.c_double(); dm = np.zeros(7)
    lib().orc_manip(_p(f64(q)), C.byref(m), _p(dm))
    return m.value, dm


def sim_time_step(x, u, ts):
    o = np.zeros(NX)
    lib().orc_sim_time_step(_p(f64(x)), _p(f64(u)), C.c_double(ts), _p(o))
    return o


def rk4(x, u, ts):
    o = np.zeros(NX)
    lib().orc_rk4(_p(f64(x)), _p(f64(u)), C.c_double(ts), _p(o))
    return o


def lin_model(Ts):
    A, B, g = np.zeros((9, 9)), np.zeros((9, 8)), np.zeros(9)
    lib().orc_lin_model(C.c_double(Ts), _p(A), _p(B), _p(g))
    return A, B, g


def rbf(delta, h):
    return lib().orc_rbf(C.c_double(delta), C.c_double(h))


def log_exp(R):
    L, E = np.zeros((3, 3)), np.zeros((3, 3))
    lib().orc_log_exp(_p(f64(R)), _p(L), _p(E))
    return L, E


def cubic_spline(x, y, xq, regular=False):
    x, y, xq = f64(x), f64(y), f64(xq)
    yq, dy, ddy = np.zeros(len(xq)), np.zeros(len(xq)), np.zeros(len(xq))
    lib().orc_cubic_spline(len(x), _p(x), _p(y), int(regular), len(xq), _p(xq), _p(yq), _p(dy), _p(ddy))
    return yq, dy, ddy


Q_HOME = np.array([0, 0, 0, -np.pi / 2, 0, np.pi / 2, np.pi / 4])
