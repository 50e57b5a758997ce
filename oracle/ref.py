"""ctypes binding of oracle/_ref/libimpec_ref.so: the REFERENCE's own translation units (pw359/Impec2D src/*.cpp, compiled
UNMODIFIED by oracle/Makefile against the shim headers in oracle/ref_shim/) behind the flat C interface of oracle/ref_api.cpp.

TEST INFRASTRUCTURE ONLY: used by tests/ (to pin the portable oracle, oracle/impec_oracle.c, to the reference itself), by
tests/golden/make_ref_golden.py (reference-generated golden vectors) and by the --impl reference leg of bench.py.  The product package
(impec2d_b200) never imports this.  The library is built in the container that holds /root/reference and travels to the GPU box as a
prebuilt file; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess
import tempfile
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}
REF_SRC = "/root/reference/src/MGSolver.cpp"


class HypreRecord(C.Structure):
    """refhypre_record (oracle/ref_shim/HYPRE_struct_ls.h): the last system the reference handed to the HYPRE struct interface."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nentries", C.c_int), ("offsets", (C.c_int * 2) * 9),
                ("matrix", C.POINTER(C.c_double)), ("rhs", C.POINTER(C.c_double)), ("solution", C.POINTER(C.c_double)),
                ("tol", C.c_double), ("iterations", C.c_int), ("final_relres", C.c_double), ("symmetric", C.c_int), ("nsolves", C.c_int)]


SOLVE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(HypreRecord), C.POINTER(C.c_double))


def so_path(omp=False):
    return os.path.join(_HERE, "_ref", "libimpec_ref_omp.so" if omp else "libimpec_ref.so")


def exe_path():
    return os.path.join(_HERE, "_ref", "Impec2D_ref")


def available(omp=False):
    """True when the compiled reference is present (built here from /root/reference, or shipped prebuilt to the GPU box)."""
    if os.path.exists(so_path(omp)):
        return True
    if os.path.exists(REF_SRC):
        try:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        except Exception:
            return False
        return os.path.exists(so_path(omp))
    return False


def lib(omp=False):
    """omp=True: the build with the reference's own OpenMP switches on (timing); False: the serial build (parity)."""
    if omp not in _LIBS:
        if not available(omp):
            raise RuntimeError("oracle/_ref is not built (needs /root/reference; run `make -C oracle ref`)")
        L = C.CDLL(so_path(omp))
        d, i, vp = C.c_double, C.c_int, C.c_void_p
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.ref_open.argtypes = [C.c_char_p]; L.ref_open.restype = vp
        L.ref_close.argtypes = [vp]
        L.ref_dims.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
        L.ref_params.argtypes = [vp, dp]
        L.ref_nparams.restype = i
        L.ref_set_threads.argtypes = [i]
        for n in ("c", "p", "v", "q"):
            getattr(L, "ref_set_" + n).argtypes = [vp, dp]
            getattr(L, "ref_get_" + n).argtypes = [vp, dp]
        L.ref_init_source.argtypes = [vp, d]; L.ref_init_source.restype = i
        L.ref_fill_boundaries.argtypes = [vp]
        L.ref_fill_neumann_velocity.argtypes = [vp]
        L.ref_mu.argtypes = [vp, d]; L.ref_mu.restype = d
        L.ref_face_mob_x.argtypes = [vp, d, d, i, i]; L.ref_face_mob_x.restype = d
        L.ref_face_mob_y.argtypes = [vp, d, d, i, i]; L.ref_face_mob_y.restype = d
        L.ref_vertex_mob_11.argtypes = [vp, i, i, d, d, d, d]; L.ref_vertex_mob_11.restype = d
        L.ref_vertex_mob_22.argtypes = [vp, i, i, d, d, d, d]; L.ref_vertex_mob_22.restype = d
        L.ref_cell_coefficients.argtypes = [vp, d, i, i, dp]
        L.ref_pressure_solve.argtypes = [vp]; L.ref_pressure_solve.restype = i
        L.ref_calc_velocities.argtypes = [vp]
        for n in ("ref_max_stable_dt", "ref_max_stable_dt_combined", "ref_max_stable_dt_advection"):
            getattr(L, n).argtypes = [vp]; getattr(L, n).restype = d
        for n in ("ref_transport_step", "ref_advect_split", "ref_advect_x", "ref_advect_y", "ref_source_split"):
            getattr(L, n).argtypes = [vp, d, i]; getattr(L, n).restype = i
        L.ref_total_mass_outflux.argtypes = [vp, i]; L.ref_total_mass_outflux.restype = d
        L.ref_flux_F.argtypes = [vp, d, i, dp]
        L.ref_flux_G.argtypes = [vp, d, i, dp]
        L.ref_diffusion_split.argtypes = [vp, d]; L.ref_diffusion_split.restype = i
        L.ref_impes_step.argtypes = [vp, d, d, C.POINTER(i)]; L.ref_impes_step.restype = d
        L.ref_profile.argtypes = [vp, i, d, d]; L.ref_profile.restype = d
        L.ref_exact_vol_frac_circle.argtypes = [d] * 5; L.ref_exact_vol_frac_circle.restype = d
        L.ref_project.argtypes = [vp, i, i, i, dp]
        L.refhypre_last.restype = C.POINTER(HypreRecord)
        L.refhypre_set_solver.argtypes = [SOLVE_FN]
        _LIBS[omp] = L
    return _LIBS[omp]


# ---------------------------------------------------------------------------------------------------------------- cfg files
# Every key the reference's SimParams reads (src/SimParams.cpp:57-432); defaults = examples/parallel.cfg.
DEFAULTS = dict(
    general=dict(starttime=0., endtime=.4, x_lb=0., y_lb=0., x_ub=1.41421356237, y_ub=1.41421356237, source_type="parallel", x_cells=141,
                 y_cells=141, num_ghostcells=2, I=1., log_to_file=False, log_skip_steps=30, log_dir="/tmp", time_matching=False,
                 log_counter_list=[], log_time_list=[], check_mass_conservation=True, max_mass_frac_error=0.0, log_mass_conservation=True,
                 check_noflow_boundary=True, check_divergence=True, results_dir="/tmp", is_test=False, gnuplot=False, levelset=False,
                 syncmode="sync"),
    elliptic=dict(eps=1e-14, ell_method="mg", bound_type="neumann", const_bound_val=0., mu_res=41., mu_inj=1., average_procedure="harmonic"),
    wells=dict(well_rc=0.05, well_profile="circle", mode="constant", q_rand_max=0.001, q_high_multiple=50., front_const_width=3,
               front_rand_width=1, well_width=0.05, q_freq=20.),
    parabolic=dict(para_method="crank-nicolson", eps=1e-14),
    hyperbolic=dict(inj_fluid_conc=1., hyper_method="muscl", dimensional_splitting="parallel", cfl=0.5, dt_diff_limiter=0.9, porosity=1.,
                    time_integration="fe", ti_eps=1e-15, nu=0.66666666666666666, diffusion=False, alpha_t=1., alpha_l=1.),
    test=dict(component="elliptic", testno=4, initial_x_cells=20, initial_y_cells=30, profile="sine", hyp_norm=1, levels=1,
              num_result_file="/tmp/ref_num.out", ex_result_file="/tmp/ref_ex.out", norm_file="/tmp/ref_norm.out",
              num_u_file="/tmp/ref_u.out", num_v_file="/tmp/ref_v.out", ex_u_file="/tmp/ref_exu.out", ex_v_file="/tmp/ref_exv.out"),
)


def _fmt(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, float):
        return repr(v) if ("." in repr(v) or "e" in repr(v) or "inf" in repr(v)) else repr(v) + "."
    if isinstance(v, str):
        return '"%s"' % v
    if isinstance(v, (list, tuple)):
        return "[" + ", ".join(_fmt(x) for x in v) + "]"
    raise TypeError(type(v))


def cfg_text(**groups):
    """libconfig text with DEFAULTS overridden per group: cfg_text(general=dict(x_cells=64), hyperbolic=dict(nu=1.))."""
    out = ["application:", "{"]
    for gname, gd in DEFAULTS.items():
        d = dict(gd)
        d.update(groups.get(gname, {}))
        out.append("%s:\n    {" % gname)
        for k, v in d.items():
            out.append("        %s = %s;" % (k, _fmt(v)))
        out.append("    };")
    unknown = set(groups) - set(DEFAULTS)
    if unknown:
        raise KeyError(unknown)
    out.append("};")
    return "\n".join(out) + "\n"


def repr_float(x):
    return repr(float(x))


class Ref:
    """One set of the reference's solver objects (SimParams, MGSolver, MUSCLSplitSolver, UpwindSplitSolver) and its arrays."""

    def __init__(self, omp=False, **groups):
        self.L = lib(omp)
        fd, self.cfg_path = tempfile.mkstemp(suffix=".cfg", prefix="impec_ref_")
        with os.fdopen(fd, "w") as f:
            f.write(cfg_text(**groups))
        self.h = self.L.ref_open(self.cfg_path.encode())
        n1, n2, g = C.c_int(), C.c_int(), C.c_int()
        self.L.ref_dims(self.h, C.byref(n1), C.byref(n2), C.byref(g))
        self.N1, self.N2, self.g = n1.value, n2.value, g.value

    def close(self):
        if self.h:
            self.L.ref_close(self.h)
            self.h = None
            try:
                os.unlink(self.cfg_path)
            except OSError:
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # shapes
    def scalar_shape(self): return (self.N1 + 2 * self.g, self.N2 + 2 * self.g)
    def velocity_shape(self): return (self.N1 + 1 + 2 * self.g, self.N2 + 1 + 2 * self.g, 2)

    def params(self):
        out = np.zeros(self.L.ref_nparams())
        self.L.ref_params(self.h, out)
        names = ["dx", "dy", "x_lb", "x_ub", "y_lb", "y_ub", "mu_res", "mu_inj", "nu", "cfl", "porosity", "eps", "para_eps", "alpha_l", "alpha_t",
                 "dt_diff_limit", "inj_fluid_conc", "I", "well_rc", "tstart", "tfinal", "stype", "avgproc", "dim_splitting", "diffusion", "ptype",
                 "is_test", "well_profile", "hmtype"]
        return dict(zip(names, out))

    def set(self, name, a): getattr(self.L, "ref_set_" + name)(self.h, np.ascontiguousarray(a, dtype=np.float64))

    def get(self, name):
        shape = {"c": self.scalar_shape(), "p": self.scalar_shape(), "v": self.velocity_shape(), "q": (self.N1 * self.N2,)}[name]
        a = np.zeros(shape)
        getattr(self.L, "ref_get_" + name)(self.h, a)
        return a

    def get_uv(self):
        v = self.get("v")
        return np.ascontiguousarray(v[:, :, 0]), np.ascontiguousarray(v[:, :, 1])

    def set_uv(self, u, v): self.set("v", np.stack([u, v], axis=-1))

    def last_system(self):
        """(matrix[N1,N2,9] in the reference's entry order C,W,SW,S,SE,E,NE,N,NW; rhs[N1,N2]; solution[N1,N2]; record)."""
        r = self.L.refhypre_last().contents
        n = r.nx * r.ny
        A = np.ctypeslib.as_array(r.matrix, shape=(n * r.nentries,)).reshape(r.nx, r.ny, r.nentries).copy()
        b = np.ctypeslib.as_array(r.rhs, shape=(n,)).reshape(r.nx, r.ny).copy()
        x = np.ctypeslib.as_array(r.solution, shape=(n,)).reshape(r.nx, r.ny).copy()
        return A, b, x, r


def run_executable(cfg_text_or_path, cwd=None, timeout=600, threads=None):
    """Run the reference's own main() (src/Driver.cpp, compiled unmodified) on a cfg; returns its stdout."""
    if not available():
        raise RuntimeError("oracle/_ref is not built")
    path = cfg_text_or_path
    tmp = None
    if not os.path.exists(path):
        fd, tmp = tempfile.mkstemp(suffix=".cfg", prefix="impec_ref_run_")
        with os.fdopen(fd, "w") as f:
            f.write(cfg_text_or_path)
        path = tmp
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    try:
        return subprocess.run([exe_path(), path], cwd=cwd, env=env, timeout=timeout, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT).stdout.decode()
    finally:
        if tmp:
            os.unlink(tmp)
