"""ctypes binding of include/mpcc_cuda.h.  No numerics here: every call goes to libmpcc_b200.so."""
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
import os as _os
LIB_PATH = Path(_os.environ.get("MPCC_B200_LIB", PKG / "libmpcc_b200.so"))   # the override serves kernel-variant experiments (tools/)
ASSETS = PKG / "assets"

NX, NU, HZ, RB, LIN = 9, 8, 17, 150, 212
PARAMS_DOUBLES, TRACK_DOUBLES = 94, 2704

EXPORTS = [
    "mpcc_cuda_last_error", "mpcc_cuda_create", "mpcc_cuda_destroy", "mpcc_cuda_upload_nn", "mpcc_cuda_load_nn", "mpcc_cuda_set_params",
    "mpcc_load_params_json", "mpcc_fit_track", "mpcc_load_track_json", "mpcc_cuda_set_tracks", "mpcc_cuda_reset", "mpcc_cuda_run_cycle",
    "mpcc_cuda_run_cycle_device", "mpcc_cuda_read_results", "mpcc_cuda_result_pointers", "mpcc_cuda_stream", "mpcc_cuda_synchronize",
    "mpcc_cuda_get_warm_state", "mpcc_cuda_set_warm_state", "mpcc_cuda_sim_time_step", "mpcc_cuda_eval_robot_data", "mpcc_cuda_eval_stage",
    "mpcc_cuda_eval_track", "mpcc_cuda_solve_ocp", "mpcc_cuda_get_stats", "mpcc_cuda_sim_time_step_device", "mpcc_cuda_set_profiling",
    "mpcc_cuda_get_kernel_times", "mpcc_cuda_fp64_peak", "mpcc_cuda_read_decisions", "mpcc_fit_tracks", "mpcc_cuda_read_qp_counters", "mpcc_cuda_read_compute_time",
    "mpcc_cuda_fit_tracks", "mpcc_cuda_get_tracks", "mpcc_track_from_knots",
    "mpcc_cuda_comm_unique_id", "mpcc_cuda_comm_init", "mpcc_cuda_gather_results", "mpcc_cuda_read_gathered", "mpcc_cuda_gathered_pointer", "mpcc_cuda_launch_count",
]


class Config(C.Structure):
    _fields_ = [("batch", C.c_int32), ("horizon", C.c_int32), ("Ts", C.c_double), ("device", C.c_int32), ("qp_max_iter", C.c_int32),
                ("qp_eps", C.c_double), ("sqp_kernel", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def lib():
    """Load libmpcc_b200.so; fails loudly when it is missing (there is no fallback path)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.mpcc_cuda_last_error.restype = C.c_char_p
        _lib.mpcc_cuda_stream.restype = C.c_void_p
        if hasattr(_lib, "mpcc_cuda_launch_count"):   # (absent only in older builds loaded through MPCC_B200_LIB for comparisons)
            _lib.mpcc_cuda_launch_count.restype = C.c_int64
            _lib.mpcc_cuda_launch_count.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"mpcc_cuda error {rc}: {lib().mpcc_cuda_last_error().decode()}")


def default_assets():
    return ASSETS


def load_default_params(param_dir=None, overrides=None):
    """Params/*.json (reference schema) -> flat record of PARAMS_DOUBLES doubles, via the library's JSON loader."""
    d = Path(param_dir) if param_dir else ASSETS / "params"
    out = np.zeros(PARAMS_DOUBLES)
    ov = overrides or {}
    keys = (C.c_char_p * max(1, len(ov)))(*[k.encode() for k in ov])
    vals = _f64(list(ov.values()) or [0.0])
    _check(lib().mpcc_load_params_json(str(d / "model.json").encode(), str(d / "cost.json").encode(), str(d / "bounds.json").encode(),
                                       str(d / "normalization.json").encode(), str(d / "sqp.json").encode(), keys, _p(vals), len(ov), _p(out)))
    return out


def fit_track(X, Y, Z, R):
    """ArcLengthSpline::gen6DSpline on the host side of the library -> spline table."""
    X, Y, Z, R = _f64(X), _f64(Y), _f64(Z), _f64(R)
    t = np.zeros(TRACK_DOUBLES)
    _check(lib().mpcc_fit_track(len(X), _p(X), _p(Y), _p(Z), _p(R), _p(t)))
    return t


def fit_tracks(X, Y, Z, R, n_threads=0):
    """Bulk ArcLengthSpline::gen6DSpline: X, Y, Z [n_tracks][n], R [n_tracks][n][9] -> tables [n_tracks][TRACK_DOUBLES]."""
    X, Y, Z, R = _f64(X), _f64(Y), _f64(Z), _f64(R)
    nt, n = X.shape
    t = np.zeros((nt, TRACK_DOUBLES))
    _check(lib().mpcc_fit_tracks(nt, n, _p(X), _p(Y), _p(Z), _p(R), _p(t), n_threads))
    return t


def track_from_knots(s, X, Y, Z, R):
    """Table of an already fitted ArcLengthSpline from its 100 knots (no fit / resample pass)."""
    t = np.zeros(TRACK_DOUBLES)
    s, X, Y, Z, R = _f64(s), _f64(X), _f64(Y), _f64(Z), _f64(R)
    assert len(s) == 100 and R.size == 900
    _check(lib().mpcc_track_from_knots(_p(s), _p(X), _p(Y), _p(Z), _p(R), _p(t)))
    return t


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it; the application hands it to every rank)."""
    buf = np.zeros(128, np.uint8)
    _check(lib().mpcc_cuda_comm_unique_id(_p(buf)))
    return buf


def load_track_json(path=None, init_position=None):
    t = np.zeros(TRACK_DOUBLES)
    ip = None if init_position is None else _f64(init_position)
    _check(lib().mpcc_load_track_json(str(path or ASSETS / "params" / "track.json").encode(), _p(ip), _p(t)))
    return t


def fp64_peak(device=0):
    """Measured FP64 FMA throughput of the device in TFLOP/s (roofline denominator)."""
    t = C.c_double()
    _check(lib().mpcc_cuda_fp64_peak(device, C.byref(t)))
    return t.value


class BatchMPC:
    """Batch of independent mpcc::MPC objects on one GPU (reference cpp/include/MPC/mpc.h:58-101)."""

    def __init__(self, batch, horizon=10, Ts=0.01, device=0, qp_max_iter=0, qp_eps=0.0, sqp_kernel=0, flags=0):
        self.B, self.N, self.S, self.Ts = int(batch), int(horizon), int(horizon) + 1, float(Ts)
        cfg = Config(self.B, self.N, self.Ts, device, qp_max_iter, qp_eps, sqp_kernel, flags)  # flags -> mpcc_cuda_config.reserved
        self.h = C.c_void_p()
        _check(lib().mpcc_cuda_create(C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().mpcc_cuda_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ----
    def load_nn(self, self_path=None, env_path=None):
        _check(lib().mpcc_cuda_load_nn(self.h, str(self_path or ASSETS / "nn" / "self_collision.f64").encode(),
                                       str(env_path or ASSETS / "nn" / "env_collision.f64").encode()))

    def set_params(self, params):
        params = _f64(params)
        n = 1 if params.ndim == 1 else params.shape[0]
        _check(lib().mpcc_cuda_set_params(self.h, _p(params), n))

    def set_tracks(self, tables, track_of_instance=None):
        tables = _f64(tables)
        n = 1 if tables.ndim == 1 else tables.shape[0]
        ids = None if track_of_instance is None else np.ascontiguousarray(track_of_instance, dtype=np.int32)
        _check(lib().mpcc_cuda_set_tracks(self.h, _p(tables), n, _p(ids)))

    def fit_tracks_device(self, X, Y, Z, R, track_of_instance=None):
        """ArcLengthSpline::fitSpline on the device for n_tracks tracks (X, Y, Z [n_tracks][n], R [n_tracks][n][9]); installs them."""
        X, Y, Z, R = _f64(X), _f64(Y), _f64(Z), _f64(R)
        if X.ndim == 1:
            X, Y, Z, R = X[None], Y[None], Z[None], R[None]
        nt, n = X.shape
        ids = None if track_of_instance is None else np.ascontiguousarray(track_of_instance, dtype=np.int32)
        _check(lib().mpcc_cuda_fit_tracks(self.h, nt, n, _p(X), _p(Y), _p(Z), _p(R), _p(ids)))

    def get_tracks(self, n_tracks=1):
        t = np.zeros((n_tracks, TRACK_DOUBLES))
        _check(lib().mpcc_cuda_get_tracks(self.h, _p(t), n_tracks))
        return t

    # ---- multi-GPU: one process / handle per GPU, gather of the per-instance results over NCCL (side stream) ----
    def comm_init(self, unique_id, rank, world):
        self.world = int(world)
        uid = None if unique_id is None else np.ascontiguousarray(unique_id, dtype=np.uint8)
        _check(lib().mpcc_cuda_comm_init(self.h, _p(uid), int(rank), int(world)))

    def gather_results(self):
        """Enqueue (non-blocking) the all-gather of this cycle's [u0 | status | iters] of every rank."""
        _check(lib().mpcc_cuda_gather_results(self.h))

    def read_gathered(self):
        n = self.B * self.world
        u = np.zeros((n, NU)); st = np.zeros(n, np.int32); it = np.zeros(n, np.int32)
        _check(lib().mpcc_cuda_read_gathered(self.h, _p(u), _p(st), _p(it)))
        return dict(u0=u, status=st, iters=it)

    def setup_default(self, init_position=None, params=None):
        self.load_nn()
        self.set_params(load_default_params() if params is None else params)
        self.set_tracks(load_track_json(None, init_position))

    def reset(self):
        _check(lib().mpcc_cuda_reset(self.h))

    # ---- the control cycle ----
    def run_cycle(self, x0, u0, obs=None, want_horizon=True, horizon_out=None):
        """horizon_out: optional caller-owned C-contiguous float64 array [B][N+1][17] for MPCReturn::mpc_horizon.  If it lives in pinned host
        memory (e.g. torch.empty(...).pin_memory().numpy()) the SQP kernel writes it directly; any other buffer is filled by a copy."""
        x0 = _f64(x0).reshape(self.B, NX).copy()
        u0 = _f64(u0).reshape(self.B, NU)
        obs = None if obs is None else _f64(obs).reshape(self.B, 4)
        u = np.zeros((self.B, NU)); hor = np.zeros((self.B, self.S, HZ)) if want_horizon else None
        if horizon_out is not None:
            assert horizon_out.dtype == np.float64 and horizon_out.flags["C_CONTIGUOUS"] and horizon_out.size == self.B * self.S * HZ
            hor = horizon_out
        st = np.zeros(self.B, np.int32); it = np.zeros(self.B, np.int32); ok = np.zeros(self.B, np.int32)
        _check(lib().mpcc_cuda_run_cycle(self.h, _p(x0), _p(u0), _p(obs), _p(u), _p(hor), _p(st), _p(it), _p(ok)))
        return dict(x0=x0, u0=u, horizon=hor, status=st, iters=it, ok=ok)

    def run_cycle_device(self, d_x0, d_u0, d_obs=None):
        """Device pointers (ints); enqueues on the handle's stream without synchronising."""
        _check(lib().mpcc_cuda_run_cycle_device(self.h, C.c_void_p(d_x0), C.c_void_p(d_u0), C.c_void_p(d_obs) if d_obs else None))

    def read_results(self, want_horizon=False):
        u = np.zeros((self.B, NU)); hor = np.zeros((self.B, self.S, HZ)) if want_horizon else None
        st = np.zeros(self.B, np.int32); it = np.zeros(self.B, np.int32); ok = np.zeros(self.B, np.int32)
        _check(lib().mpcc_cuda_read_results(self.h, _p(u), _p(hor), _p(st), _p(it), _p(ok)))
        return dict(u0=u, horizon=hor, status=st, iters=it, ok=ok)

    def result_pointers(self):
        ptrs = [C.c_void_p() for _ in range(5)]
        _check(lib().mpcc_cuda_result_pointers(self.h, *[C.byref(p) for p in ptrs]))
        return [p.value for p in ptrs]

    @property
    def stream(self):
        return lib().mpcc_cuda_stream(self.h)

    def synchronize(self):
        _check(lib().mpcc_cuda_synchronize(self.h))

    def get_warm_state(self):
        hor = np.zeros((self.B, self.S, HZ)); v = np.zeros(self.B, np.int32); f = np.zeros(self.B, np.int32)
        _check(lib().mpcc_cuda_get_warm_state(self.h, _p(hor), _p(v), _p(f)))
        return hor, v, f

    def set_warm_state(self, horizon, valid, failed):
        hor = _f64(horizon).reshape(self.B, self.S, HZ)
        v = np.ascontiguousarray(valid, dtype=np.int32); f = np.ascontiguousarray(failed, dtype=np.int32)
        _check(lib().mpcc_cuda_set_warm_state(self.h, _p(hor), _p(v), _p(f)))

    def sim_time_step(self, x, u, ts=None):
        x = _f64(x).reshape(self.B, NX); u = _f64(u).reshape(self.B, NU); xn = np.zeros((self.B, NX))
        _check(lib().mpcc_cuda_sim_time_step(self.h, _p(x), _p(u), C.c_double(self.Ts if ts is None else ts), _p(xn)))
        return xn

    def sim_time_step_device(self, d_x, d_u, d_xn, ts=None):
        _check(lib().mpcc_cuda_sim_time_step_device(self.h, C.c_void_p(d_x), C.c_void_p(d_u), C.c_double(self.Ts if ts is None else ts), C.c_void_p(d_xn)))

    def set_profiling(self, on=True):
        _check(lib().mpcc_cuda_set_profiling(self.h, int(on)))

    def kernel_times(self):
        """ms of [prologue, kinematics, networks, SQP] of the last cycle (profiling must be on)."""
        t = np.zeros(4)
        _check(lib().mpcc_cuda_get_kernel_times(self.h, _p(t)))
        return t

    # ---- per-function evaluators ----
    def eval_robot_data(self, q, obs=None):
        q = _f64(q).reshape(-1, 7); n = q.shape[0]
        obs = None if obs is None else _f64(obs).reshape(n, 4)
        rb = np.zeros((n, RB))
        _check(lib().mpcc_cuda_eval_robot_data(self.h, _p(q), _p(obs), n, _p(rb)))
        return rb

    def eval_stage(self, x, u, u_prev, u_next, x_next, rb, k):
        x = _f64(x).reshape(-1, NX); n = x.shape[0]
        out = np.zeros((n, LIN))
        kk = np.ascontiguousarray(k, dtype=np.int32)
        _check(lib().mpcc_cuda_eval_stage(self.h, _p(x), _p(_f64(u)), _p(_f64(u_prev)), _p(_f64(u_next)), _p(_f64(x_next)), _p(_f64(rb)), _p(kk), n, _p(out)))
        return out

    def eval_track(self, s):
        s = _f64(s).ravel(); out = np.zeros((len(s), 21))
        _check(lib().mpcc_cuda_eval_track(self.h, _p(s), len(s), _p(out)))
        return out

    def solve_ocp(self, guess, rb, cur_u, max_log=0, want_steps=False):
        g = _f64(guess).reshape(-1, self.S, HZ).copy(); n = g.shape[0]
        st = np.zeros(n, np.int32); it = np.zeros(n, np.int32); nl = np.zeros(n, np.int32)
        al = np.zeros((n, max(1, max_log))); steps = np.zeros((n, max(1, max_log), self.S, HZ)) if want_steps else None
        _check(lib().mpcc_cuda_solve_ocp(self.h, _p(g), _p(_f64(rb)), _p(_f64(cur_u)), n, _p(st), _p(it), _p(steps), _p(al), max_log, _p(nl)))
        return dict(horizon=g, status=st, iters=it, alphas=al, n_logged=nl, steps=steps)

    def compute_time(self):
        """[B][4] seconds: total, set_qp, solve_qp, get_alpha of the last cycle's solveOCP (per instance)."""
        t = np.zeros((self.B, 4))
        _check(lib().mpcc_cuda_read_compute_time(self.h, _p(t)))
        return t

    def qp_counters(self):
        a = np.zeros(self.B, np.int32); b = np.zeros(self.B, np.int32)
        _check(lib().mpcc_cuda_read_qp_counters(self.h, _p(a), _p(b)))
        return a, b

    def decisions(self):
        m = np.zeros(self.B, np.int32)
        _check(lib().mpcc_cuda_read_decisions(self.h, _p(m)))
        return m.view(np.uint32)

    def launch_count(self):
        """kernels launched by the last cycle (host-side counter, no synchronisation)"""
        return int(lib().mpcc_cuda_launch_count(self.h))

    def stats(self):
        s = np.zeros(6, np.int64)
        _check(lib().mpcc_cuda_get_stats(self.h, _p(s)))
        return dict(launches=int(s[0]), sqp_iters=int(s[1]), qp_iters=int(s[2]), qp_fail=int(s[3]), solved=int(s[4]), ok=int(s[5]))
