"""CPU tier: the C-ABI shared library loads, exports every symbol include/mpcc_cuda.h declares, its host-side
loaders (JSON params, track fit; no device needed) agree with the oracle, and -- without a GPU -- the compute
entry points fail loudly instead of falling back to anything."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from helpers import flat_params, f64

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    txt = (ROOT / "include" / "mpcc_cuda.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcc_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from mpcc_manipulator_b200 import capi
    lib = capi.lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(capi.EXPORTS) == names


def test_params_json_loader_matches_oracle(O):
    from mpcc_manipulator_b200 import capi
    assert np.array_equal(capi.load_default_params(), flat_params(O.load_params()))
    ov = capi.load_default_params(overrides={"cost.qC": 750.0, "model.tol_sing": 0.018, "sqp.eps_prim": 1e-3})
    ref = flat_params(O.load_params(overrides={"cost": {"qC": 750.0}, "model": {"tol_sing": 0.018}, "sqp": {"eps_prim": 1e-3}}))
    # the solver interface keeps the cost FILE's rddq (osqp_interface.cpp:28,57): last entry unaffected by overrides
    assert np.array_equal(ov, ref)
    with pytest.raises(RuntimeError):
        capi.load_default_params(param_dir="/nonexistent")
    with pytest.raises(RuntimeError):
        capi.load_default_params(overrides={"nofile.key": 1.0})


def test_track_fit_host_matches_oracle(O, nn, track_wp):
    from mpcc_manipulator_b200 import capi
    t = capi.fit_track(*track_wp)
    o = O.OracleMPC(N=10, nn=nn)
    o.set_track(*track_wp)
    s, X, Y, Z, R = o.track_table()
    assert np.abs(t[:100] - s).max() < 1e-12
    assert np.abs(t[100:200] - X).max() < 1e-12  # a coefficients of X = knot values
    ee = O.fk(O.Q_HOME)[0]
    t2 = capi.load_track_json(None, ee)
    assert np.array_equal(t, t2)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mpcc_manipulator_b200 as M
    with pytest.raises(RuntimeError, match="CUDA"):
        M.BatchMPC(4, 10)


def test_create_argument_validation():
    from mpcc_manipulator_b200 import capi
    lib = capi.lib()
    h = C.c_void_p()
    for cfg in (capi.Config(0, 10, 0.01, 0, 0, 0.0, 0, 0), capi.Config(4, 1, 0.01, 0, 0, 0.0, 0, 0), capi.Config(4, 10, -1.0, 0, 0, 0.0, 0, 0)):
        assert lib.mpcc_cuda_create(C.byref(cfg), C.byref(h)) == 1  # MPCC_ERR_INVALID
        assert len(lib.mpcc_cuda_last_error()) > 0
    assert lib.mpcc_cuda_create(None, C.byref(h)) == 1


def test_bulk_track_fit_matches_single_fits():
    from mpcc_manipulator_b200 import capi
    rng = np.random.default_rng(2)
    nt, n = 37, 100
    t = np.linspace(np.pi / 2, 5 * np.pi / 2, n)
    a, b, c = rng.uniform(1.5, 3, nt), rng.uniform(1.5, 3, nt), rng.uniform(0, 2.5, nt)
    X = 0.1 * a[:, None] * np.sin(t)[None] + 0.55; Y = 0.1 * b[:, None] * np.sin(2 * t)[None]; Z = 0.1 * c[:, None] * np.cos(t)[None] + 0.52
    R = np.tile(np.diag([1., -1., -1.]).ravel(), (nt, n, 1))
    tabs = capi.fit_tracks(X, Y, Z, R, n_threads=4)
    for i in (0, 5, 36):
        assert np.array_equal(tabs[i], capi.fit_track(X[i], Y[i], Z[i], R[i]))


def test_device_track_fit_code_matches_host_fit(O):
    """csrc/dev_track_fit.cuh (the one-thread-per-track kernel's code) compiled for the host, contiguous and strided scratch
    (the kernel's [element][track] layout), against the host fit that the oracle pins: tracks of the track.py family
    (cpp/Params/track.py:5-22) with 5 .. 160 waypoints and varying orientation."""
    import ctypes as C
    from helpers import Emul, _p
    emu = Emul()
    rng = np.random.default_rng(1)
    ee = O.fk(O.Q_HOME)[0]
    worst = 0.0
    for trial in range(12):
        n = int(rng.integers(5, 160))
        a, bb = rng.uniform(1.5, 3, 2); c = rng.uniform(0, 2.5)
        t = np.linspace(np.pi / 2, 5 * np.pi / 2, n)
        X, Y, Z = O.shift_track(a * 0.1 * np.sin(t), bb * 0.1 * np.sin(2 * t), c * 0.1 * np.cos(t), ee)
        R = []
        for i in range(n):
            w = np.array([0.3 * np.sin(t[i]), 0.2 * np.cos(t[i]), 0.1 * t[i]]) * (trial % 2)
            th = np.linalg.norm(w)
            K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            E = np.eye(3) + (np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K if th > 0 else 0)
            R.append((np.diag([1., -1., -1.]) @ E).ravel())
        R = np.array(R)
        ref = emu.fit_track(X, Y, Z, R)
        for stride in (1, 5):
            out = np.zeros_like(ref)
            emu.lib.emu_fit_track_device_code(n, _p(f64(X)), _p(f64(Y)), _p(f64(Z)), _p(f64(R)), stride, _p(out))
            worst = max(worst, np.abs(out - ref).max() / np.abs(ref).max())
    assert worst < 1e-12, worst


def test_track_from_knots_reproduces_the_fitted_table(track_wp):
    """mpcc_track_from_knots: the 100 knots of a fitted ArcLengthSpline give back the same table without re-fitting
    (the binding of SolverInterface::setTrack(ArcLengthSpline), solver_interface.h:46)."""
    from mpcc_manipulator_b200 import capi
    t = capi.fit_track(*track_wp)
    s, X, Y, Z = t[:100], t[100:200], t[200:300], t[300:400]      # knots s_i and the a-coefficients (= knot values) of X, Y, Z
    R = t[1300:2200]
    t2 = capi.track_from_knots(s, X, Y, Z, R)
    assert np.abs(t2 - t).max() < 1e-13 * np.abs(t).max()
    with pytest.raises(RuntimeError, match="increase"):
        capi.track_from_knots(s[::-1].copy(), X, Y, Z, R)
