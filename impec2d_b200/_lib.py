"""ctypes loader for the C-ABI library (include/impec2d_b200.h).

The product library is impec2d_b200/libimpec2d_b200.so (hand-written sm_100a kernels).  There is no CPU
fallback: if the library is missing or no CUDA device is present, calls fail loudly.  The CPU-emulated
build (libimpec2d_emu.so) is test infrastructure and is only loaded when a test asks for it by name.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


class Params(C.Structure):
    """impec_params: the subset of the reference's SimParams the hot path reads (src/SimParams.h:29-173)."""
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


SENDRECV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong)
BCAST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int)


class CommCallbacks(C.Structure):
    _fields_ = [("user", C.c_void_p), ("sendrecv", SENDRECV_FN), ("allreduce", ALLREDUCE_FN), ("allgather", ALLGATHER_FN), ("bcast", BCAST_FN)]


class Diag(C.Structure):
    _fields_ = [("total_mass", C.c_double), ("divergence_error", C.c_double), ("boundary_velocity_error", C.c_double),
                ("total_mass_outflux", C.c_double), ("c_min", C.c_double), ("c_max", C.c_double)]


class IntegralDiag(C.Structure):
    _fields_ = [("produced_fraction", C.c_double), ("mixing_zone_length", C.c_double), ("mixing_zone_cells", C.c_double), ("mass_in_place", C.c_double)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("relres", C.c_double), ("true_relres", C.c_double), ("converged", C.c_int)]


EXPORTS = [
    "impec_default_params", "impec_version", "impec_is_emulated", "impec_nccl_unique_id", "impec_ctx_create",
    "impec_ctx_destroy", "impec_last_error", "impec_synchronize", "impec_upload_c", "impec_upload_p", "impec_upload_q",
    "impec_upload_velocity", "impec_download_c", "impec_download_p", "impec_download_velocity", "impec_field_layout",
    "impec_field_dev", "impec_set_q_dev_changed", "impec_fill_boundaries", "impec_pressure_solve", "impec_calc_velocities",
    "impec_max_stable_dt", "impec_source_split", "impec_advect_x", "impec_advect_y", "impec_advect_split", "impec_intercell_flux", "impec_gather_rows",
    "impec_diffusion_split", "impec_transport_step", "impec_impes_step", "impec_diagnostics", "impec_integral_diagnostics", "impec_build_operator",
    "impec_vcycle", "impec_mg_levels", "impec_mg_w_range", "impec_kernel_launches", "impec_overlapped_exchanges", "impec_last_parabolic_info", "impec_reset_state", "impec_set_option",
    "impec_profile_kernel", "impec_init_source", "impec_init_horizontal_front", "impec_synth_state", "impec_ctx_set_comm_callbacks", "impec_local_rows",
    "impec_upload_c_rows", "impec_upload_q_rows", "impec_download_c_rows", "impec_init_source_rows", "impec_synth_state_rows",
    "impec_comm_backend", "impec_transfer_bytes", "impec_bind_host_to_device", "impec_project_dims", "impec_project_to_diagonal", "impec_checkpoint_save", "impec_checkpoint_load",
]


def library_path(emulated=False):
    return os.path.join(_HERE, "libimpec2d_emu.so" if emulated else "libimpec2d_b200.so")


def _preload_nccl():
    """libimpec2d_b200.so needs libnccl.so.2.  PyTorch ships its own, newer NCCL (nvidia/nccl/lib) and libtorch_cuda.so fails to
    load (`undefined symbol: ncclDevCommCreate`) if the system's older libnccl.so.2 is already in the process.  So whichever of
    torch / this library is imported first, make the bundled NCCL the one that gets loaded: processes that use both (bench.py,
    the multi-GPU tests) then share one NCCL.  No torch import here, and no error if the wheel is absent (system NCCL is used)."""
    import sys
    for base in sys.path:
        cand = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
            return


def load(emulated=False):
    """Load the C-ABI library.  emulated=True is for tests only (CPU emulation of the CUDA kernels)."""
    key = bool(emulated)
    if key in _CACHE:
        return _CACHE[key]
    path = library_path(emulated)
    if not os.path.exists(path):
        raise RuntimeError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % path)
    if not emulated:
        _preload_nccl()
    L = C.CDLL(path)
    if bool(L.impec_is_emulated()) != key:
        raise RuntimeError("library %s reports emulated=%d, expected %d" % (path, L.impec_is_emulated(), key))
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    dptr = C.POINTER(C.c_double)
    L.impec_default_params.argtypes = [C.POINTER(Params)]
    L.impec_version.restype = C.c_char_p
    L.impec_nccl_unique_id.argtypes = [C.c_char_p]
    L.impec_bind_host_to_device.argtypes = [C.c_int, C.c_char_p, C.c_int]
    L.impec_ctx_create.argtypes = [C.POINTER(Params), i, i, i, C.c_char_p, vp, C.POINTER(vp)]
    L.impec_ctx_destroy.argtypes = [vp]; L.impec_ctx_destroy.restype = None
    L.impec_last_error.argtypes = [vp]; L.impec_last_error.restype = C.c_char_p
    L.impec_comm_backend.argtypes = [vp]; L.impec_comm_backend.restype = C.c_char_p
    L.impec_synchronize.argtypes = [vp]
    for n in ("impec_upload_c", "impec_upload_p", "impec_upload_q", "impec_upload_velocity",
              "impec_download_c", "impec_download_p", "impec_download_velocity"):
        getattr(L, n).argtypes = [vp, vp]
    L.impec_field_layout.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.impec_field_dev.argtypes = [vp, C.c_char_p]; L.impec_field_dev.restype = vp
    for n in ("impec_set_q_dev_changed", "impec_fill_boundaries", "impec_calc_velocities", "impec_build_operator", "impec_reset_state"):
        getattr(L, n).argtypes = [vp]
    L.impec_pressure_solve.argtypes = [vp, d, C.POINTER(SolveInfo)]
    L.impec_max_stable_dt.argtypes = [vp, dptr]
    for n in ("impec_source_split", "impec_advect_x", "impec_advect_y", "impec_advect_split", "impec_transport_step"):
        getattr(L, n).argtypes = [vp, d]
    L.impec_diffusion_split.argtypes = [vp, d, C.POINTER(SolveInfo)]
    L.impec_impes_step.argtypes = [vp, d, d, dptr, C.POINTER(SolveInfo)]
    L.impec_diagnostics.argtypes = [vp, C.POINTER(Diag)]
    L.impec_integral_diagnostics.argtypes = [vp, d, d, C.POINTER(IntegralDiag)]
    L.impec_project_dims.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.impec_project_to_diagonal.argtypes = [vp, i, vp]
    L.impec_checkpoint_save.argtypes = [vp, C.c_char_p, d, C.c_longlong]
    L.impec_checkpoint_load.argtypes = [vp, C.c_char_p, dptr, C.POINTER(C.c_longlong)]
    L.impec_vcycle.argtypes = [vp, vp, vp]
    L.impec_mg_levels.argtypes = [vp]
    L.impec_mg_w_range.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.impec_intercell_flux.argtypes = [vp, i, d, vp]
    L.impec_gather_rows.argtypes = [vp, C.c_char_p, vp]
    L.impec_kernel_launches.argtypes = [vp]; L.impec_kernel_launches.restype = C.c_longlong
    L.impec_last_parabolic_info.argtypes = [vp, C.POINTER(SolveInfo)]
    L.impec_overlapped_exchanges.argtypes = [vp]; L.impec_overlapped_exchanges.restype = C.c_longlong
    L.impec_set_option.argtypes = [vp, C.c_char_p, d]
    L.impec_profile_kernel.argtypes = [vp, C.c_char_p, i, dptr]
    L.impec_init_source.argtypes = [C.POINTER(Params), d, d, i, vp]
    L.impec_init_horizontal_front.argtypes = [C.POINTER(Params), d, d, i, d, d, d, d, d, vp]
    L.impec_synth_state.argtypes = [C.POINTER(Params), C.c_ulonglong, vp]
    L.impec_ctx_set_comm_callbacks.argtypes = [vp, C.POINTER(CommCallbacks)]
    L.impec_local_rows.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.impec_transfer_bytes.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    for n in ("impec_upload_c_rows", "impec_upload_q_rows", "impec_download_c_rows"):
        getattr(L, n).argtypes = [vp, vp]
    L.impec_init_source_rows.argtypes = [C.POINTER(Params), d, d, i, i, i, vp]
    L.impec_synth_state_rows.argtypes = [C.POINTER(Params), C.c_ulonglong, i, i, vp]
    _CACHE[key] = L
    return L
