"""B200-native batched MPCC control cycle (the hot path of JunHeonYoon/MPCC_manipulator).

The product is the CUDA shared library ``libmpcc_b200.so`` (C ABI: ``include/mpcc_cuda.h``) and the C++
host classes under ``csrc/host``.  This Python package is a thin ctypes binding used by the tests and
``bench.py``; it contains no numerical code and has no CPU fallback: importing :mod:`.capi` raises if the
library has not been built (``python -c "import __graft_entry__ as g; g.build()"``).
"""
from .capi import BatchMPC, load_default_params, fit_track, fit_tracks, load_track_json, track_from_knots, comm_unique_id, default_assets, fp64_peak, LIB_PATH  # noqa: F401
