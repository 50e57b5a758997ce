"""Shared case builders for the parity tests.  `emulated=True` runs the SAME kernel source through the CPU
emulator (tests only); `emulated=False` calls the real sm_100a library through the C ABI."""
import ctypes as C
import numpy as np
import oracle as O
import impec2d_b200 as I

SQRT2 = 1.41421356237      # examples/parallel.cfg:19-20


def both_params(emulated, **kw):
    return O.default_params(**kw), I.default_params(emulated=emulated, **kw)


def make_context(Pi, emulated):
    return I.Context(Pi, emulated=emulated)


def production_case(N1, N2=None, stype=O.PARALLEL, seed=20260917, **kw):
    """parallel.cfg / diagonal_dispersion.cfg style inputs on an N1 x N2 grid: developed synthetic
    concentration, quarter-circle wells."""
    N2 = N1 if N2 is None else N2
    L = SQRT2 if stype == O.PARALLEL else 1.0
    base = dict(N1=N1, N2=N2, x_ub=L, y_ub=L, stype=stype)
    base.update(kw)
    Po = O.default_params(**base)
    lib = O.lib()
    c = O.scalar_field(Po)
    lib.orc_synth_state(C.byref(Po), seed, c)
    q = np.zeros(N1 * N2)
    if stype == O.HORIZONTAL:      # source_type "horizontal": injection layer at i = 0.., production layer at the far end (src/Util.cpp:607-639)
        assert lib.orc_init_horizontal_front(C.byref(Po), 1.0, 0.05, 2, 1.0, 1.0, 0.0, 1.0, 0.0, q) >= 1
    else:
        lib.orc_init_source(C.byref(Po), 1.0, 0.05, 0, q)
    return base, Po, c, q


def advection_case(N, profile=O.SINE, **kw):
    """hyperbolic_test.cfg: periodic advection with (u,v) = (-Lx, +Ly) (src/Impec2DSolver.cpp:1003-1015)."""
    base = dict(N1=N, N2=N, x_ub=1.0, y_ub=1.0, is_test=1, testno=1, stype=O.DIAGONAL, cfl=0.95, diffusion=0)
    base.update(kw)
    Po = O.default_params(**base)
    lib = O.lib()
    c = O.scalar_field(Po)
    g = Po.g
    for i in range(N):
        for j in range(N):
            c[g + i, g + j] = lib.orc_profile(C.byref(Po), profile, Po.x_lb + (i + 0.5) * Po.dx, Po.y_lb + (j + 0.5) * Po.dy)
    u = O.velocity_field(Po, -(Po.x_ub - Po.x_lb))
    v = O.velocity_field(Po, +(Po.y_ub - Po.y_lb))
    return base, Po, c, u, v


def aos(u, v):
    return np.ascontiguousarray(np.stack([u, v], axis=-1))


def interior(P, a):
    g = P.g
    return a[g:g + P.N1, g:g + P.N2]


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
