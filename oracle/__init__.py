"""ctypes binding of the CPU ORACLE (oracle/impec_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package (impec2d_b200) never imports this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PARALLEL, DIAGONAL, HORIZONTAL = 0, 1, 2
MEANCONC, HARMONIC = 0, 1
ALTERNATING, X_FIRST, Y_FIRST, STRANG, SYMMETRIC_PARALLEL = 0, 1, 2, 3, 4
MUSCL, UPWIND = 0, 1
EXPLICIT, CRANKNICOLSON = 0, 1
SINE, TOPHAT, GAUSS, HSINE = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [
        ("N1", C.c_int), ("N2", C.c_int), ("g", C.c_int),
        ("x_lb", C.c_double), ("x_ub", C.c_double), ("y_lb", C.c_double), ("y_ub", C.c_double),
        ("mu_res", C.c_double), ("mu_inj", C.c_double), ("avgproc", C.c_int),
        ("nu", C.c_double), ("is_test", C.c_int), ("testno", C.c_int), ("stype", C.c_int),
        ("porosity", C.c_double), ("cfl", C.c_double), ("inj_fluid_conc", C.c_double),
        ("diffusion", C.c_int), ("alpha_l", C.c_double), ("alpha_t", C.c_double), ("dt_diff_limit", C.c_double),
        ("ptype", C.c_int), ("eps", C.c_double), ("para_eps", C.c_double),
        ("hyper_method", C.c_int), ("dim_splitting", C.c_int),
        ("max_iter", C.c_int), ("mg_max_levels", C.c_int), ("mg_nmin", C.c_int), ("mg_ncoarse", C.c_int),
        ("mg_w_from", C.c_int), ("mg_w_to", C.c_int),
    ]

    @property
    def dx(self):
        return (self.x_ub - self.x_lb) / self.N1

    @property
    def dy(self):
        return (self.y_ub - self.y_lb) / self.N2


class State(C.Structure):
    _fields_ = [("total_mass_outflux", C.c_double), ("prev_first_sweep", C.c_int),
                ("last_its", C.c_int), ("last_parabolic_its", C.c_int),
                ("last_relres", C.c_double), ("last_true_relres", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("impec_oracle.c", "impec_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        d = C.c_double
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        pp = C.POINTER(Params)
        sp = C.POINTER(State)
        L = _LIB
        L.orc_default_params.argtypes = [pp]
        L.orc_state_init.argtypes = [sp]
        for name in ("orc_fill_periodic_scalar", "orc_fill_neumann_scalar"):
            getattr(L, name).argtypes = [dp, C.c_int, C.c_int, C.c_int]
        L.orc_fill_neumann_velocity.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int]
        L.orc_fill_boundaries.argtypes = [pp, dp]
        L.orc_mu.argtypes = [pp, d]; L.orc_mu.restype = d
        L.orc_assemble_elliptic.argtypes = [pp, dp, dp, dp, dp]
        L.orc_matvec9.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.orc_solve_elliptic.argtypes = [pp, sp, dp, dp, dp]; L.orc_solve_elliptic.restype = C.c_int
        L.orc_mgpcg.argtypes = [pp, C.c_int, C.c_int, dp, dp, C.c_int, d, dp, C.POINTER(C.c_int), C.POINTER(d), C.POINTER(d)]
        L.orc_mgpcg.restype = C.c_int
        L.orc_mg_apply.argtypes = [pp, C.c_int, C.c_int, dp, C.c_int, dp, dp]; L.orc_mg_apply.restype = C.c_int
        L.orc_calc_velocities.argtypes = [pp, dp, dp, dp, dp]
        L.orc_project_dims.argtypes = [pp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_project_parallel.argtypes = [pp, dp, C.c_int, C.c_int, dp]
        for name in ("orc_max_stable_dt", "orc_max_stable_advection_dt", "orc_max_stable_dt_combined"):
            getattr(L, name).argtypes = [pp, dp, dp]; getattr(L, name).restype = d
        L.orc_tr_integrate.argtypes = [pp, dp, dp, d]; L.orc_tr_integrate.restype = d
        L.orc_flux_F.argtypes = [pp, C.c_int, dp, dp, d, dp]
        L.orc_flux_G.argtypes = [pp, C.c_int, dp, dp, d, dp]
        L.orc_advect_x.argtypes = [pp, dp, dp, d]
        L.orc_advect_y.argtypes = [pp, dp, dp, d]
        L.orc_advect_split.argtypes = [pp, sp, dp, dp, dp, d]
        L.orc_assemble_parabolic.argtypes = [pp, dp, dp, dp, d, dp, dp]
        L.orc_solve_parabolic.argtypes = [pp, sp, dp, dp, dp, d]; L.orc_solve_parabolic.restype = C.c_int
        L.orc_transport_step.argtypes = [pp, sp, dp, dp, dp, dp, d]; L.orc_transport_step.restype = C.c_int
        L.orc_impes_step.argtypes = [pp, sp, dp, dp, dp, dp, dp, d, d]; L.orc_impes_step.restype = d
        L.orc_exact_vol_frac_circle.argtypes = [d] * 5; L.orc_exact_vol_frac_circle.restype = d
        L.orc_init_source.argtypes = [pp, d, d, C.c_int, dp]
        L.orc_init_horizontal_front.argtypes = [pp, d, d, C.c_int, d, d, d, d, d, dp]; L.orc_init_horizontal_front.restype = C.c_int
        L.orc_profile.argtypes = [pp, C.c_int, d, d]; L.orc_profile.restype = d
        L.orc_advection_norm.argtypes = [pp, dp, C.c_int, C.c_int]; L.orc_advection_norm.restype = d
        L.orc_init_elliptic_test.argtypes = [pp, dp, dp, dp, dp]
        L.orc_total_mass.argtypes = [pp, dp]; L.orc_total_mass.restype = d
        L.orc_divergence_error.argtypes = [pp, dp, dp, dp]; L.orc_divergence_error.restype = d
        L.orc_boundary_velocity_error.argtypes = [pp, dp, dp]; L.orc_boundary_velocity_error.restype = d
        L.orc_synth_state.argtypes = [pp, C.c_ulonglong, dp]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
    return _LIB


def default_params(**kw):
    P = Params()
    lib().orc_default_params(C.byref(P))
    for k, v in kw.items():
        if not hasattr(P, k):
            raise AttributeError(k)
        setattr(P, k, v)
    return P


def new_state():
    S = State()
    lib().orc_state_init(C.byref(S))
    return S


def scalar_field(P, fill=0.0):
    return np.full((P.N1 + 2 * P.g, P.N2 + 2 * P.g), fill, dtype=np.float64)


def velocity_field(P, fill=0.0):
    return np.full((P.N1 + 1 + 2 * P.g, P.N2 + 1 + 2 * P.g), fill, dtype=np.float64)


def interior(P, a):
    g = P.g
    return a[g:g + P.N1, g:g + P.N2]
