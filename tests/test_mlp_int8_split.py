"""CPU tier: arithmetic of the int8-split ("Ozaki") layers of the MLP kernel (csrc/mlp_oz_kernel.cuh).

The three 256 x 256 layers of the env net run on tcgen05 as s8 x s8 -> s32 products of 7-bit digit planes.  What the tensor
cores do is exact integer arithmetic, so everything that decides the result -- the host packing of the weight digits, the
column scales, the digit extraction, the int64 Horner recombination -- is plain scalar code, shared verbatim between the
kernel and tests/emul (oz_col_scales, oz_quantize, oz_digit, oz_wq_index, pack_mlp_oz_weights).  Here that code is run on the
host with the products restated as int32 sums and compared with a float64 / exact-rational product; the GPU tier repeats the
comparison on the device (test_gpu_parity.py::test_robot_data_int8_split_vs_fp64_kernel and the RobotData tests against the
oracle).  Bar: 1e-9 relative for the linearisation inputs (north_star); the split is held to 1e-12 of full scale here."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

from helpers import Emul, f64

_p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


@pytest.fixture(scope="module")
def emu():
    return Emul()


def oz_layer(emu, W, X):
    W, X = f64(W), f64(X)
    Y = np.zeros((256, X.shape[1]))
    md = C.c_int(0)
    mg = C.c_longlong(0)
    emu.lib.emu_oz_layer(_p(W), _p(X), C.c_int(X.shape[1]), _p(Y), C.byref(md), C.byref(mg))
    return Y, md.value, mg.value


def test_split_layer_matches_fp64_product(emu):
    rng = np.random.default_rng(11)
    S = emu.lib.emu_oz_slices()
    assert 5 <= S <= 7
    W = rng.normal(0, 0.08, (256, 256))
    # activations as the network produces them: half the neurons masked to exactly 0, columns of very different magnitude
    X = rng.normal(0, 1.0, (256, 16)) * (rng.random((256, 16)) < 0.5)
    X *= 10.0 ** rng.uniform(-6, 3, 16)[None, :]
    Y, md, mg = oz_layer(emu, W, X)
    ref = W @ X
    full = np.abs(W).max(axis=1)[:, None] * np.abs(X).max(axis=0)[None, :] * 256  # |W|_max |X|_max K: the scale of the dropped terms
    assert md <= 64, md                      # signed 7-bit digits
    assert mg < 2 ** 23, mg                  # int32 accumulators are far from overflow
    err = np.abs(Y - ref) / full
    assert err.max() < (S + 1) * 2.0 ** (-7 * S) + 1e-15, err.max()
    # and against the entries themselves (a well-conditioned random product): far inside the 1e-9 bar of the linearisation inputs
    assert np.abs(Y - ref).max() / np.abs(ref).max() < 1e-12


def test_split_layer_exact_rational_reference(emu):
    """One column against an exact rational dot product: the split's error is the dropped digit products only."""
    rng = np.random.default_rng(12)
    S = emu.lib.emu_oz_slices()
    W = rng.normal(0, 0.1, (256, 256))
    X = rng.normal(0, 3.0, (256, 1))
    Y, _, _ = oz_layer(emu, W, X)
    for r in (0, 17, 255):
        exact = sum(Fraction(float(W[r, k])) * Fraction(float(X[k, 0])) for k in range(256))
        full = float(np.abs(W[r]).max() * np.abs(X).max() * 256)
        assert abs(float(Fraction(float(Y[r, 0])) - exact)) / full < (S + 1) * 2.0 ** (-7 * S) + 2e-16


def test_split_layer_edge_columns(emu):
    """All-zero columns, a single nonzero entry, entries at the column maximum (top digit +-64), tiny and huge columns, exact powers of two."""
    rng = np.random.default_rng(13)
    W = rng.normal(0, 0.05, (256, 256))
    W[3, :] = 0.0                      # a zero weight row
    W[5, 7] = 2.0 ** -3                # exact power of two as the row maximum
    W[5, :7] = 0.0; W[5, 8:] = 0.0
    X = np.zeros((256, 8))
    X[10, 1] = 1.0
    X[:, 2] = rng.choice([-1.0, 1.0], 256) * 0.75      # every entry at the column maximum
    X[:, 3] = rng.normal(0, 1, 256) * 1e-200            # far below any scale the network produces, still above the zero threshold
    X[:, 4] = rng.normal(0, 1, 256) * 1e150
    X[:, 5] = 2.0 ** rng.integers(-20, 0, 256)          # powers of two spread over 20 binades
    X[:, 6] = rng.normal(0, 1, 256) * 1e-300            # below the threshold (biased exponent < 7 S): treated as a zero column
    Y, md, mg = oz_layer(emu, W, X)
    ref = W @ X
    assert md <= 64 and mg < 2 ** 23
    assert np.all(Y[:, 0] == 0.0) and np.all(Y[3, :] == 0.0)
    assert np.all(Y[:, 6] == 0.0)
    for c in (1, 2, 3, 4, 5):
        full = np.abs(W).max(axis=1) * np.abs(X[:, c]).max() * 256
        ok = full > 0
        assert (np.abs(Y[ok, c] - ref[ok, c]) / full[ok]).max() < 1e-13, c
    assert Y[5, 1] == 0.0 and abs(Y[5, 5] - ref[5, 5]) <= 2.0 ** -52 * abs(ref[5, 5])


def test_relu_mask_decisions_survive_the_split(emu):
    """Pre-activations close to zero: the sign (= the ReLU mask) of W x + b must agree with float64 unless |W x + b| is inside the split's own
    error bound -- the same statement test_relu_mask_agreement makes for the fp64 kernel's summation order."""
    rng = np.random.default_rng(14)
    S = emu.lib.emu_oz_slices()
    W = rng.normal(0, 0.08, (256, 256))
    X = np.abs(rng.normal(0, 1.0, (256, 32))) * (rng.random((256, 32)) < 0.5)
    ref = W @ X
    b = -ref + rng.normal(0, 1, ref.shape) * 10.0 ** rng.uniform(-13, -3, ref.shape) * np.abs(ref).max()   # biases that put W x + b near zero
    Y, _, _ = oz_layer(emu, W, X)
    full = np.abs(W).max(axis=1)[:, None] * np.abs(X).max(axis=0)[None, :] * 256
    bound = ((S + 1) * 2.0 ** (-7 * S) + 256 * 2.0 ** -53) * full   # dropped digit products + the float64 reference's own summation error
    pre_ref, pre_oz = ref + b, Y + b
    disagree = (pre_ref > 0) != (pre_oz > 0)
    assert np.all(np.abs(pre_ref[disagree]) <= bound[disagree])
    assert disagree.sum() < 0.01 * disagree.size   # ties at the 1e-14 level only, even in this adversarial set
