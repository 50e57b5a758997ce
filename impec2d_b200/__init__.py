"""impec2d_b200 -- B200-native IMPES hot path for pw359/Impec2D.

The product is the C-ABI shared library built from impec2d_b200/csrc (hand-written sm_100a CUDA,
include/impec2d_b200.h) and the C++ host classes in impec2d_b200/host that mirror the reference's
operator surface.  This Python package is the thin ctypes view of the same C ABI that the tests and
bench.py drive; it adds no numerics of its own.
"""
from ._lib import Params, Diag, SolveInfo, load, library_path, EXPORTS  # noqa: F401
from .api import (Context, default_params, ImpecError, init_source, init_horizontal_front, synth_state, init_source_rows, synth_state_rows,  # noqa: F401
                  attach_torch_comm, bind_host_to_device)

PARALLEL, DIAGONAL, HORIZONTAL = 0, 1, 2
MEANCONC, HARMONIC = 0, 1
ALTERNATING, X_FIRST, Y_FIRST, STRANG, SYMMETRIC_PARALLEL = 0, 1, 2, 3, 4
MUSCL, UPWIND = 0, 1
EXPLICIT, CRANKNICOLSON = 0, 1
