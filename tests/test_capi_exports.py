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
