"""Python view of the C ABI: one method per entry point of include/impec2d_b200.h."""
import ctypes as C
import numpy as np
from . import _lib


class ImpecError(RuntimeError):
    pass


def default_params(emulated=False, **kw):
    P = _lib.Params()
    _lib.load(emulated).impec_default_params(C.byref(P))
    for k, v in kw.items():
        if not hasattr(P, k):
            raise AttributeError(k)
        setattr(P, k, v)
    return P


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Owns one impec_ctx (one GPU).  Host arrays use the reference's layouts:
    c, p : (N1+2g, N2+2g); velocity : (N1+1+2g, N2+1+2g, 2); q : (N1*N2,)."""

    def __init__(self, params, device=0, rank=0, nranks=1, nccl_id=None, stream=None, emulated=False):
        self.L = _lib.load(emulated)
        self.P = params
        self._h = C.c_void_p()
        rc = self.L.impec_ctx_create(C.byref(params), device, rank, nranks, nccl_id, stream, C.byref(self._h))
        if rc != 0:
            msg = self.L.impec_last_error(self._h).decode() if self._h else "context allocation failed"
            if self._h:
                self.L.impec_ctx_destroy(self._h)
                self._h = C.c_void_p()
            raise ImpecError("impec_ctx_create failed (%d): %s" % (rc, msg))

    def close(self):
        if self._h:
            self.L.impec_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc > 0:          # IMPEC_WARN_NOCONV: the call completed with the last iterate (upstream never checks convergence)
            import warnings
            warnings.warn("impec: %s" % self.L.impec_last_error(self._h).decode())
            return
        if rc != 0:
            raise ImpecError("impec call failed (%d): %s" % (rc, self.L.impec_last_error(self._h).decode()))

    # shapes
    def scalar_shape(self):
        return (self.P.N1 + 2 * self.P.g, self.P.N2 + 2 * self.P.g)

    def velocity_shape(self):
        return (self.P.N1 + 1 + 2 * self.P.g, self.P.N2 + 1 + 2 * self.P.g, 2)

    # transfers
    def upload_c(self, c): assert c.shape == self.scalar_shape(); self._chk(self.L.impec_upload_c(self._h, _ptr(c)))
    def upload_p(self, p): assert p.shape == self.scalar_shape(); self._chk(self.L.impec_upload_p(self._h, _ptr(p)))
    def upload_q(self, q): assert q.size == self.P.N1 * self.P.N2; self._chk(self.L.impec_upload_q(self._h, _ptr(q)))
    def upload_velocity(self, v): assert v.shape == self.velocity_shape(); self._chk(self.L.impec_upload_velocity(self._h, _ptr(v)))

    def download_c(self, out=None):
        out = np.empty(self.scalar_shape()) if out is None else out
        self._chk(self.L.impec_download_c(self._h, _ptr(out)))
        return out

    def download_p(self, out=None):
        out = np.empty(self.scalar_shape()) if out is None else out
        self._chk(self.L.impec_download_p(self._h, _ptr(out)))
        return out

    def download_velocity(self, out=None):
        out = np.empty(self.velocity_shape()) if out is None else out
        self._chk(self.L.impec_download_velocity(self._h, _ptr(out)))
        return out

    # path
    def fill_boundaries(self): self._chk(self.L.impec_fill_boundaries(self._h))

    def pressure_solve(self, eps=None):
        info = _lib.SolveInfo()
        self._chk(self.L.impec_pressure_solve(self._h, self.P.eps if eps is None else eps, C.byref(info)))
        return info

    def calc_velocities(self): self._chk(self.L.impec_calc_velocities(self._h))

    def max_stable_dt(self):
        dt = C.c_double()
        self._chk(self.L.impec_max_stable_dt(self._h, C.byref(dt)))
        return dt.value

    def source_split(self, dt): self._chk(self.L.impec_source_split(self._h, dt))
    def advect_x(self, dt): self._chk(self.L.impec_advect_x(self._h, dt))
    def advect_y(self, dt): self._chk(self.L.impec_advect_y(self._h, dt))
    def advect_split(self, dt): self._chk(self.L.impec_advect_split(self._h, dt))

    def diffusion_split(self, dt):
        info = _lib.SolveInfo()
        self._chk(self.L.impec_diffusion_split(self._h, dt, C.byref(info)))
        return info

    def transport_step(self, dt): self._chk(self.L.impec_transport_step(self._h, dt))

    def gather_rows(self, field):
        """Whole-grid host copy of a slab-decomposed field ("c", "p": ghosted; "v": AoS velocity); collective over the ranks."""
        P = self.P; g = P.g
        out = np.zeros((P.N1 + 2 * g, P.N2 + 2 * g)) if field in ("c", "p") else np.zeros((P.N1 + 1 + 2 * g, P.N2 + 1 + 2 * g, 2))
        self._chk(self.L.impec_gather_rows(self._h, field.encode(), _ptr(out)))
        return out

    def intercell_flux(self, direction, dt):
        """CalculateIntercellFlux_F (direction 0, shape (N1+1, N2+2g)) / _G (direction 1, shape (N1+2g, N2+1)) of the current c and velocity."""
        P = self.P; g = P.g
        out = np.zeros((P.N1 + 1, P.N2 + 2 * g) if direction == 0 else (P.N1 + 2 * g, P.N2 + 1))
        self._chk(self.L.impec_intercell_flux(self._h, direction, dt, _ptr(out)))
        return out

    def impes_step(self, t=0.0, t_next_output=0.0):
        dt = C.c_double(); info = _lib.SolveInfo()
        self._chk(self.L.impec_impes_step(self._h, t, t_next_output, C.byref(dt), C.byref(info)))
        return dt.value, info

    def diagnostics(self):
        d = _lib.Diag()
        self._chk(self.L.impec_diagnostics(self._h, C.byref(d)))
        return d

    def integral_diagnostics(self, lo=0.1, hi=0.9):
        """produced fraction, mixing-zone length, mass in place as device reductions (no field download)."""
        d = _lib.IntegralDiag()
        self._chk(self.L.impec_integral_diagnostics(self._h, lo, hi, C.byref(d)))
        return d

    def project_to_diagonal(self, field=0):
        """Impec2DSolver::Project: c (field 0) or p (field 1) of a parallel run on the diagonal grid (dN1 x dN2 array)."""
        n1 = C.c_int(); n2 = C.c_int()
        self._chk(self.L.impec_project_dims(self._h, C.byref(n1), C.byref(n2)))
        out = np.empty((n1.value, n2.value))
        self._chk(self.L.impec_project_to_diagonal(self._h, field, out.ctypes.data_as(C.c_void_p)))
        return out

    def checkpoint_save(self, path, t=0.0, step=0): self._chk(self.L.impec_checkpoint_save(self._h, str(path).encode(), t, step))

    def checkpoint_load(self, path):
        t = C.c_double(); step = C.c_longlong()
        self._chk(self.L.impec_checkpoint_load(self._h, str(path).encode(), C.byref(t), C.byref(step)))
        return t.value, step.value

    def build_operator(self): self._chk(self.L.impec_build_operator(self._h))

    def vcycle(self, r=None):
        if r is None:
            self._chk(self.L.impec_vcycle(self._h, None, None))
            return None
        z = np.empty_like(r)
        self._chk(self.L.impec_vcycle(self._h, _ptr(r), _ptr(z)))
        return z

    def mg_levels(self): return self.L.impec_mg_levels(self._h)

    def mg_w_range(self):
        a, b = C.c_int(), C.c_int()
        self._chk(self.L.impec_mg_w_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value
    def kernel_launches(self): return self.L.impec_kernel_launches(self._h)
    def last_parabolic_info(self):
        info = _lib.SolveInfo(); self._chk(self.L.impec_last_parabolic_info(self._h, C.byref(info))); return info

    def overlapped_exchanges(self): return self.L.impec_overlapped_exchanges(self._h)
    def reset_state(self): self._chk(self.L.impec_reset_state(self._h))
    def set_option(self, key, value): self._chk(self.L.impec_set_option(self._h, key.encode(), float(value)))
    def synchronize(self): self._chk(self.L.impec_synchronize(self._h))
    def comm_backend(self): return self.L.impec_comm_backend(self._h).decode()

    def transfer_bytes(self):
        h = C.c_ulonglong(); d = C.c_ulonglong()
        self._chk(self.L.impec_transfer_bytes(self._h, C.byref(h), C.byref(d)))
        return h.value, d.value

    def bytes_h2d(self): return self.transfer_bytes()[0]
    def bytes_d2h(self): return self.transfer_bytes()[1]

    def local_rows(self):
        r0 = C.c_int(); n = C.c_int()
        self._chk(self.L.impec_local_rows(self._h, C.byref(r0), C.byref(n)))
        return r0.value, n.value

    def upload_c_rows(self, c): self._chk(self.L.impec_upload_c_rows(self._h, _ptr(c)))
    def upload_q_rows(self, q): self._chk(self.L.impec_upload_q_rows(self._h, _ptr(q)))

    def download_c_rows(self, out):
        self._chk(self.L.impec_download_c_rows(self._h, _ptr(out)))
        return out

    def profile_kernel(self, name, reps=10):
        ms = C.c_double()
        self._chk(self.L.impec_profile_kernel(self._h, name.encode(), reps, C.byref(ms)))
        return ms.value


def init_horizontal_front(params, I=1.0, well_width=0.05, mode=2, q_freq=1.0, q_high_multiple=1.0, tstart=0.0, tfinal=1.0, t=0.0, emulated=False):
    """InitHorizontalFront (host): q[N1*N2] of source_type "horizontal"; mode 0 pulsing, 1 linear, 2 constant."""
    q = np.zeros(params.N1 * params.N2)
    rc = _lib.load(emulated).impec_init_horizontal_front(C.byref(params), I, well_width, mode, q_freq, q_high_multiple, tstart, tfinal, t, _ptr(q))
    if rc != 0:
        raise ImpecError("impec_init_horizontal_front failed (%d): the injection layer must be 1 .. N1/2 rows wide" % rc)
    return q


def init_source(params, I=1.0, well_rc=0.05, well_profile=0, emulated=False):
    """Impec2DSolver::InitializeSource (host): returns q[N1*N2]."""
    q = np.zeros(params.N1 * params.N2)
    rc = _lib.load(emulated).impec_init_source(C.byref(params), I, well_rc, well_profile, _ptr(q))
    if rc != 0:
        raise ImpecError("impec_init_source failed (%d)" % rc)
    return q


def init_source_rows(params, row0, nrows, I=1.0, well_rc=0.05, well_profile=0, emulated=False):
    q = np.zeros(nrows * params.N2)
    rc = _lib.load(emulated).impec_init_source_rows(C.byref(params), I, well_rc, well_profile, row0, nrows, _ptr(q))
    if rc != 0:
        raise ImpecError("impec_init_source_rows failed (%d)" % rc)
    return q


def synth_state_rows(params, row0, nrows, seed=20260917, out=None, emulated=False):
    c = np.zeros((nrows + 2 * params.g, params.N2 + 2 * params.g)) if out is None else out
    rc = _lib.load(emulated).impec_synth_state_rows(C.byref(params), seed, row0, nrows, _ptr(c))
    if rc != 0:
        raise ImpecError("impec_synth_state_rows failed (%d)" % rc)
    return c


def synth_state(params, seed=20260917, out=None, emulated=False):
    """Developed synthetic concentration (ghosted array, interior written)."""
    c = np.zeros((params.N1 + 2 * params.g, params.N2 + 2 * params.g)) if out is None else out
    rc = _lib.load(emulated).impec_synth_state(C.byref(params), seed, _ptr(c))
    if rc != 0:
        raise ImpecError("impec_synth_state failed (%d)" % rc)
    return c


def bind_host_to_device(device, emulated=False):
    """impec_bind_host_to_device: restrict this thread (and the threads it starts later) to the CPUs local to `device`.  Returns
    (number of CPUs, the kernel's cpu list) or (0, None) when the topology is not visible / not allowed to this process (mask unchanged)."""
    L = _lib.load(emulated)
    buf = C.create_string_buffer(4096)
    n = L.impec_bind_host_to_device(int(device), buf, len(buf))
    return (n, buf.value.decode()) if n > 0 else (0, None)


def attach_torch_comm(ctx, rank, nranks):
    """Tests only (CPU-emulated library): route the slab halo exchanges and small collectives of a multi-rank context
    through torch.distributed (gloo).  The CUDA library talks NCCL itself and does not use this."""
    import torch
    import torch.distributed as dist

    def view(ptr, n):
        return torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(int(n),)))

    def sendrecv(user, send_lo, recv_lo, send_hi, recv_hi, n):
        # peers modulo the rank count (NULL pointers mark walls; the periodic self test passes all four and closes the ranks to a ring);
        # one tag per direction: with two ranks both neighbours are the same peer.  Tag 1 = travelling down (send_lo -> the peer's recv_hi).
        ops = []
        lo, hi = (rank - 1) % nranks, (rank + 1) % nranks
        if send_lo:
            ops += [dist.P2POp(dist.isend, view(send_lo, n).clone(), lo, tag=1), dist.P2POp(dist.irecv, view(recv_lo, n), lo, tag=2)]
        if send_hi:
            ops += [dist.P2POp(dist.isend, view(send_hi, n).clone(), hi, tag=2), dist.P2POp(dist.irecv, view(recv_hi, n), hi, tag=1)]
        for r in dist.batch_isend_irecv(ops) if ops else []:
            r.wait()
        return 0

    def allreduce(user, buf, n, op):
        dist.all_reduce(view(buf, n), op={0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MAX, 2: dist.ReduceOp.MIN}[op])
        return 0

    def allgather(user, send, recv, n):
        out = view(recv, n * nranks)
        dist.all_gather_into_tensor(out, view(send, n).clone())
        return 0

    def bcast(user, buf, n, root):
        dist.broadcast(view(buf, n), src=root)
        return 0

    cb = _lib.CommCallbacks(None, _lib.SENDRECV_FN(sendrecv), _lib.ALLREDUCE_FN(allreduce), _lib.ALLGATHER_FN(allgather), _lib.BCAST_FN(bcast))
    ctx._comm_cb = cb                                  # keep the ctypes thunks alive
    ctx._chk(ctx.L.impec_ctx_set_comm_callbacks(ctx._h, C.byref(cb)))
