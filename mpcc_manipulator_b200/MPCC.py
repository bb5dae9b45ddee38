"""Python surface of the reference, on the CUDA path.

Mirrors ``python/MPCC/MPCC.py:10-114`` and ``python/MPCC/robot_model.py`` of JunHeonYoon/MPCC_manipulator (which wrap the
Boost.Python module ``MPCC_WRAPPER``, ``cpp/src/MPCC_wrapper.cpp:116-416``): same class names, method names, argument meaning
and return tuples, so that ``python/main.py``-style scripts run unchanged with ``from mpcc_manipulator_b200.MPCC import MPCC``.
``BatchMPCC`` is the same surface with a leading batch dimension on every array.  No numerics here: everything goes through
``libmpcc_b200.so`` (capi.BatchMPC); without a GPU the constructor raises.
"""
import json
from pathlib import Path

import numpy as np

from . import capi

NX, NU, N_DEFAULT, PANDA_DOF, PANDA_NUM_LINKS = 9, 8, 10, 7, 9   # config.h:31-38, robot_model.h:13

PARAM_DICT = {   # python/MPCC/MPCC.py:37-43
    "param": ["max_dist_proj", "desired_ee_velocity", "s_trust_region", "tol_sing", "tol_selcol", "tol_envcol"],
    "cost": ["qC", "qCNmult", "qL", "qVs", "qOri", "qSing", "rdq", "rddq", "rdVs", "qC_reduction_ratio", "qL_increase_ratio", "qOri_reduction_ratio"],
    "bounds": ["q1l", "q2l", "q3l", "q4l", "q5l", "q6l", "q7l", "sl", "vsl", "q1u", "q2u", "q3u", "q4u", "q5u", "q6u", "q7u", "su", "vsu",
               "dq1l", "dq2l", "dq3l", "dq4l", "dq5l", "dq6l", "dq7l", "dVsl", "dq1u", "dq2u", "dq3u", "dq4u", "dq5u", "dq6u", "dq7u", "dVsu"],
    "normalization": ["q1", "q2", "q3", "q4", "q5", "q6", "q7", "s", "vs", "dq1", "dq2", "dq3", "dq4", "dq5", "dq6", "dq7", "dVs"],
    "sqp": ["eps_prim", "eps_dual", "line_search_tau", "line_search_eta", "line_search_rho", "max_iter", "line_search_max_iter", "do_SOC", "use_BFGS"],
}


def _check_param_value(param_value):
    assert set(param_value.keys()).issubset(PARAM_DICT), f"List of Parameters must be a subset of {list(PARAM_DICT)}, but got {list(param_value.keys())}"
    for key, value in param_value.items():
        assert set(value.keys()).issubset(PARAM_DICT[key]), f"Keys for {key} must be a subset of {PARAM_DICT[key]}, but got {list(value.keys())}"


def live_overrides(param_value):
    """What MPC::setParam changes on a LIVE object (reference quirk 13, osqp_interface.cpp:95-100 + mpc.cpp:204-209): the model
    ("param") and cost maps take effect; `bounds_` is rebuilt from the FILE (the map is ignored), the normalisation and SQP
    parameters and the solver's own copy of rddq are never updated."""
    ov = {}
    for k, v in param_value.get("param", {}).items():
        ov[f"model.{k}"] = float(v)
    for k, v in param_value.get("cost", {}).items():
        ov[f"cost.{k}"] = float(v)
    return ov


class RobotModel:
    """python/MPCC/robot_model.py: kinematics of the Panda through the library's RobotData evaluator (GPU)."""

    def __init__(self, _mpc=None):
        self._own = _mpc is None
        self._mpc = _mpc or capi.BatchMPC(1, 2)
        if self._own:
            self._mpc.load_nn()
        self.num_q = PANDA_DOF

    def _rb(self, joint_angle):
        q = np.asarray(joint_angle, dtype=np.float64)
        assert q.shape[-1] == self.num_q, f"Joint angle size {q.shape[-1]} does not match expected size {self.num_q}"
        return self._mpc.eval_robot_data(q.reshape(-1, 7)), q.ndim == 1

    def _out(self, a, single):
        return a[0] if single else a

    def getEEJacobian(self, joint_angle):
        rb, one = self._rb(joint_angle)
        return self._out(np.concatenate([rb[:, 19:40].reshape(-1, 3, 7), rb[:, 40:61].reshape(-1, 3, 7)], axis=1), one)  # [linear; angular], robot_model.cpp:372-375

    def getEEJacobianv(self, joint_angle):
        rb, one = self._rb(joint_angle)
        return self._out(rb[:, 19:40].reshape(-1, 3, 7), one)

    def getEEJacobianw(self, joint_angle):
        rb, one = self._rb(joint_angle)
        return self._out(rb[:, 40:61].reshape(-1, 3, 7), one)

    def getEEPosition(self, joint_angle):
        rb, one = self._rb(joint_angle)
        return self._out(rb[:, 7:10], one)

    def getEEOrientation(self, joint_angle):
        rb, one = self._rb(joint_angle)
        return self._out(rb[:, 10:19].reshape(-1, 3, 3), one)

    def getEEManipulability(self, joint_angle):
        rb, one = self._rb(joint_angle)
        return self._out(rb[:, 61], one)


class _SplinePath:
    """ArcLengthSpline::getPathData (PathData: X, Y, Z, R, s at the 100 knots)."""

    def __init__(self, table):
        t = np.asarray(table)
        self.s, self.X, self.Y, self.Z = t[:100].copy(), t[100:200].copy(), t[200:300].copy(), t[300:400].copy()
        self.R = [t[1300 + 9 * i:1309 + 9 * i].reshape(3, 3).copy() for i in range(100)]
        self.n_points = 100


class BatchMPCC:
    """The MPCC surface for a batch of independent controllers on one GPU: every array argument / result carries a leading
    batch dimension.  `horizon` is the reference's compile-time N (config.h:36, default 10)."""

    def __init__(self, batch=1, horizon=N_DEFAULT, device=0, param_dir=None, param_value=None):
        self.param_dir = Path(param_dir) if param_dir else capi.ASSETS / "params"
        self.jsonConfig = json.loads((self.param_dir / "config.json").read_text())
        self.Ts = float(self.jsonConfig["Ts"])
        self.batch = int(batch)
        self.pred_horizon = int(horizon)
        self.robot_dof = PANDA_DOF
        self.num_links = PANDA_NUM_LINKS
        self.mpc = capi.BatchMPC(self.batch, self.pred_horizon, Ts=self.Ts, device=device)
        self.mpc.load_nn()
        # MPC(Ts, path, param_value) (mpc.cpp:36-52): every map but "bounds" takes effect at construction
        pv = param_value or {}
        _check_param_value(pv)
        self._ctor_over = {f"{'model' if f == 'param' else f}.{k}": float(v) for f in ("param", "cost", "normalization", "sqp") for k, v in pv.get(f, {}).items()}
        self.mpc.set_params(capi.load_default_params(self.param_dir, overrides=self._ctor_over))
        self.robot_model = RobotModel(self.mpc)
        self.track_set = False

    def close(self):
        self.mpc.close()

    def setParam(self, param_value: dict) -> None:
        _check_param_value(param_value)
        keep = {k: v for k, v in self._ctor_over.items() if k.startswith(("normalization.", "sqp."))}   # never updated on a live object
        self.mpc.set_params(capi.load_default_params(self.param_dir, overrides={**keep, **live_overrides(param_value)}))

    def setTrack(self, state) -> None:
        state = np.asarray(state, dtype=np.float64).reshape(self.batch, -1)
        assert state.shape[1] == NX, f"State size {state.shape[1]} does not match expected size {NX}"
        self.init_state = state
        ee = self.robot_model.getEEPosition(state[:, :7])
        track_path = self.param_dir / "track.json"
        tables = np.stack([capi.load_track_json(track_path, ee[b]) for b in range(self.batch)])  # Track::getTrack(ee_pos) + gen6DSpline
        self.mpc.set_tracks(tables, np.arange(self.batch))
        self.tables = tables
        self.spline_paths = [_SplinePath(t) for t in tables]
        self.track_set = True

    def getSplinePath(self):
        assert self.track_set, "Set Track first!"
        pos = np.stack([np.stack([p.X, p.Y, p.Z], axis=1) for p in self.spline_paths])
        rot = np.stack([np.array(p.R) for p in self.spline_paths])
        arc = np.stack([p.s for p in self.spline_paths])
        return pos, rot, arc

    def runMPC(self, state, input, obs_position=None, obs_radius=None):
        assert self.track_set, "Set Track first!"
        state = np.asarray(state, dtype=np.float64).reshape(self.batch, -1)
        assert state.shape[1] == NX, f"State size {state.shape[1]} does not match expected size {NX}"
        u = np.asarray(input, dtype=np.float64).reshape(self.batch, NU)
        obs = None
        if obs_position is not None:
            p = np.broadcast_to(np.asarray(obs_position, dtype=np.float64).reshape(-1, 3), (self.batch, 3))
            r = np.broadcast_to(np.asarray(0.0 if obs_radius is None else obs_radius, dtype=np.float64).reshape(-1), (self.batch,))
            obs = np.c_[p, r]
        r = self.mpc.run_cycle(state, u, obs)
        ct = self.mpc.compute_time()
        compute_time = {"total": ct[:, 0], "set_qp": ct[:, 1], "solve_qp": ct[:, 2], "get_alpha": ct[:, 3], "set_env": np.zeros(self.batch)}
        return r["ok"].astype(bool), r["x0"], r["u0"], r["horizon"], compute_time


class MPCC(BatchMPCC):
    """python/MPCC/MPCC.py:10-114 -- one controller (a batch of one)."""

    def __init__(self, horizon=N_DEFAULT, device=0, param_dir=None, param_value=None) -> None:
        super().__init__(1, horizon, device, param_dir, param_value)

    def setTrack(self, state) -> None:
        state = np.asarray(state, dtype=np.float64)
        assert state.size == NX, f"State size {state.size} does not match expected size {NX}"
        super().setTrack(state.reshape(1, NX))
        self.init_state = state
        self.spline_path = self.spline_paths[0]

    def getSplinePath(self):
        pos, rot, arc = super().getSplinePath()
        return pos[0], rot[0], arc[0]

    def getRefPose(self, path_parameter: float):
        s = self.spline_path.s
        assert path_parameter >= np.min(s) - 1E-3 and path_parameter <= np.max(s) + 1E-3, f"Path parameter must be in [{np.min(s), np.max(s)}] and your input is {path_parameter}"
        o = self.mpc.eval_track([path_parameter])[0]     # ArcLengthSpline::getPosition / getOrientation on the device
        return o[0:3].copy(), o[9:18].reshape(3, 3).copy()

    def getContourError(self, s: float, ee_posi):
        return np.linalg.norm(self.mpc.eval_track([s])[0, 0:3] - np.asarray(ee_posi))

    def runMPC(self, state, input, obs_position=np.array([3, 3, 3]), obs_radius: float = 0):
        state = np.asarray(state, dtype=np.float64)
        assert state.size == NX, f"State size {state.size} does not match expected size {NX}"
        ok, x, u, hor, ct = super().runMPC(state.reshape(1, NX), np.asarray(input, dtype=np.float64).reshape(1, NU), obs_position, obs_radius)
        mpc_horizon = [{"state": hor[0, k, :NX].copy(), "input": hor[0, k, NX:].copy()} for k in range(self.pred_horizon + 1)]
        compute_time = {k: float(v[0]) for k, v in ct.items()}
        return bool(ok[0]), x[0], u[0], mpc_horizon, compute_time
