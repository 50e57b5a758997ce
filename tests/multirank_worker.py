"""Worker for tests/test_multirank.py: one process per rank (gloo), CPU-emulated kernel library.  Runs the same
problem on nranks slabs and, on every rank, also on a single-rank context; the slab result must agree with this
rank's rows of the single-rank result."""
import os
import sys
import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import impec2d_b200 as I   # noqa: E402
import oracle as O         # noqa: E402
from helpers import production_case   # noqa: E402


def slab_context(Pi, cuda, rank, world, dev):
    """This rank's slab context: the real library over NCCL on GPUs, or the emulator over gloo."""
    if cuda:
        import ctypes as C
        ids = [None]
        if rank == 0:
            buf = C.create_string_buffer(128)
            assert I.load().impec_nccl_unique_id(buf) == 0
            ids = [buf.raw]
        dist.broadcast_object_list(ids, src=0)
        return I.Context(Pi, device=dev, rank=rank, nranks=world, nccl_id=ids[0])
    ctx = I.Context(Pi, rank=rank, nranks=world, emulated=True)
    I.attach_torch_comm(ctx, rank, world)
    return ctx


def elliptic_self_test(N1, N2, testno, cuda, rank, world):
    """elliptic_test.cfg on slabs (src/EllipticSolver.cpp:191-239): manufactured K(x,y) fields, p = cos x cos y; the slab solve against the
    1-rank solve of the same library and against the exact solution."""
    import ctypes as C
    from helpers import interior
    kw = dict(N1=N1, N2=N2, x_ub=2 * np.pi, y_ub=2 * np.pi, mu_res=1.0, mu_inj=1.0, avgproc=O.MEANCONC, nu=float(os.environ.get("IMPEC_TEST_NU", "1.0")),
              is_test=1, testno=testno, stype=O.DIAGONAL, eps=1e-12)
    Po = O.default_params(**kw); Pi = I.default_params(emulated=not cuda, **kw)
    c = O.scalar_field(Po); pex = O.scalar_field(Po); uex = O.velocity_field(Po); vex = O.velocity_field(Po); q = np.zeros(N1 * N2)
    O.lib().orc_init_elliptic_test(C.byref(Po), pex, q, uex, vex)
    dev = int(os.environ.get("LOCAL_RANK", "0")) if cuda else 0
    ref = I.Context(Pi, device=dev, emulated=not cuda)
    ref.upload_c(c); ref.upload_q(q); info_ref = ref.pressure_solve(); ref.calc_velocities()
    pr = ref.download_p(); vr = ref.download_velocity()
    ctx = slab_context(Pi, cuda, rank, world, dev)
    ctx.upload_c(c); ctx.upload_q(q); info = ctx.pressure_solve(); ctx.calc_velocities()
    assert info.converged == 1 and abs(info.iterations - info_ref.iterations) <= 1, (info.iterations, info_ref.iterations)
    pg = ctx.gather_rows("p"); vg = ctx.gather_rows("v")
    g = Pi.g
    dp = np.abs(pg[g:-g, g:-g] - pr[g:-g, g:-g]).max() / np.abs(pr).max()
    dv = max(np.abs(vg[g:g + N1, g:g + N2, 0] - vr[g:g + N1, g:g + N2, 0]).max(), np.abs(vg[g:g + N1, g:g + N2 + 1, 1] - vr[g:g + N1, g:g + N2 + 1, 1]).max())
    assert dp <= 1e-9 and dv <= 1e-8, (dp, dv)
    err = np.abs(interior(Po, pg) - interior(Po, pex)).sum() * Po.dx * Po.dy
    err_ref = np.abs(interior(Po, pr) - interior(Po, pex)).sum() * Po.dx * Po.dy
    assert abs(err - err_ref) <= 1e-6 * err_ref, (err, err_ref)
    print("rank %d ok: dp %.2e dv %.2e its %d/%d err %.3e" % (rank, dp, dv, info.iterations, info_ref.iterations, err), flush=True)
    ctx.close(); ref.close()
    dist.destroy_process_group()


def periodic_self_test(N, steps, upwind, cuda, rank, world):
    """hyperbolic_test.cfg on slabs: periodic advection with the prescribed velocity (-1, +1); the ranks are closed to a ring for the ghost
    rows.  The transport path is explicit, so the slab run equals the 1-rank run BIT FOR BIT."""
    from helpers import advection_case, aos
    kw = dict(hyper_method=O.UPWIND) if upwind else {}
    base, Po, c, u, v = advection_case(N, **kw)
    Pi = I.default_params(emulated=not cuda, **base)
    q = np.zeros(N * N)
    dev = int(os.environ.get("LOCAL_RANK", "0")) if cuda else 0
    ref = I.Context(Pi, device=dev, emulated=not cuda)
    ctx = slab_context(Pi, cuda, rank, world, dev)
    out = []
    for k in (ref, ctx):
        k.upload_c(c); k.upload_velocity(aos(u, v)); k.upload_q(q)
        dt = k.max_stable_dt()
        for _ in range(steps):
            k.transport_step(dt)
        out.append((dt, k.gather_rows("c"), k.diagnostics().total_mass))
    g = Pi.g
    assert out[0][0] == out[1][0], (out[0][0], out[1][0])
    assert np.array_equal(out[0][1][g:-g, g:-g], out[1][1][g:-g, g:-g]), np.abs(out[0][1][g:-g, g:-g] - out[1][1][g:-g, g:-g]).max()
    assert abs(out[0][2] - out[1][2]) <= 1e-12 * abs(out[0][2])
    assert np.abs(out[1][1][g:-g, g:-g] - c[g:-g, g:-g]).max() > 1e-3, "the profile did not move"
    print("rank %d ok: periodic advection on %d slabs bit-identical after %d steps" % (rank, world, steps), flush=True)
    ctx.close(); ref.close()
    dist.destroy_process_group()


def main():
    N1, N2, stype, diffusion, steps = (int(x) for x in sys.argv[1:6])
    cuda = len(sys.argv) > 6 and sys.argv[6] == "cuda"      # real library + NCCL on GPUs instead of emulator + gloo
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if os.environ.get("IMPEC_TEST_PERIODIC"):
        return periodic_self_test(N1, steps, os.environ["IMPEC_TEST_PERIODIC"] == "upwind", cuda, rank, world)
    if os.environ.get("IMPEC_TEST_ELLIPTIC"):
        return elliptic_self_test(N1, N2, int(os.environ["IMPEC_TEST_ELLIPTIC"]), cuda, rank, world)
    extra = {}
    if os.environ.get("IMPEC_TEST_SPLITTING"):      # dimensional_splitting other than the symmetric-parallel default
        extra["dim_splitting"] = int(os.environ["IMPEC_TEST_SPLITTING"])
    if os.environ.get("IMPEC_TEST_ALPHA_T"):        # alpha_l != alpha_t: non-symmetric Crank-Nicolson operator, BiCGStab
        extra["alpha_t"] = float(os.environ["IMPEC_TEST_ALPHA_T"])
    if os.environ.get("IMPEC_TEST_UPWIND"):
        extra["hyper_method"] = O.UPWIND
    base, Po, c, q = production_case(N1, N2, stype=stype, diffusion=diffusion, eps=1e-12, **extra)
    Pi = I.default_params(emulated=not cuda, **base)
    dev = int(os.environ.get("LOCAL_RANK", "0")) if cuda else 0

    ref = I.Context(Pi, device=dev, emulated=not cuda)      # single-rank reference (same library, same kernels)
    ref.upload_c(c); ref.upload_q(q)
    ctx = slab_context(Pi, cuda, rank, world, dev)
    ctx.upload_c(c); ctx.upload_q(q)
    r0, n = ctx.local_rows()
    assert n == N1 // world and r0 == rank * n
    g = Pi.g
    t = 0.0
    for s in range(steps):
        dt_ref, info_ref = ref.impes_step(t, 0.0)
        dt, info = ctx.impes_step(t, 0.0)
        assert abs(dt - dt_ref) <= 1e-10 * dt_ref, (dt, dt_ref)
        assert info.converged == 1 and abs(info.iterations - info_ref.iterations) <= 1, (info.iterations, info_ref.iterations)
        t += dt_ref
    cg = np.zeros_like(c); pg = np.zeros_like(c)
    ctx.download_c(cg); ctx.download_p(pg)
    cr = ref.download_c(); pr = ref.download_p()
    rows = slice(g + r0, g + r0 + n)
    dc = np.abs(cg[rows, g:-g] - cr[rows, g:-g]).max()
    dp = np.abs(pg[rows, g:-g] - pr[rows, g:-g]).max() / np.abs(pr).max()
    d = ctx.diagnostics(); dr = ref.diagnostics()
    assert dc <= 1e-9, dc
    assert dp <= 1e-9, dp
    assert abs(d.total_mass - dr.total_mass) <= 1e-9 * dr.total_mass
    assert abs(d.total_mass_outflux - dr.total_mass_outflux) <= 1e-9 * max(dr.total_mass_outflux, 1e-30)
    assert abs(d.c_max - dr.c_max) <= 1e-9 and abs(d.c_min - dr.c_min) <= 1e-9, (d.c_max, dr.c_max)
    assert abs(d.divergence_error - dr.divergence_error) <= 1e-9 + 1e-3 * dr.divergence_error, (d.divergence_error, dr.divergence_error)
    assert abs(d.boundary_velocity_error - dr.boundary_velocity_error) <= 1e-9, (d.boundary_velocity_error, dr.boundary_velocity_error)
    # impec_gather_rows: the whole field assembled from all slabs must equal the 1-rank field (interior rows; all columns of c / p)
    cga = ctx.gather_rows("c"); pgl = ctx.gather_rows("p"); vg = ctx.gather_rows("v"); vr = ref.download_velocity()
    assert np.abs(cga[g:-g, :] - cr[g:-g, :]).max() <= 1e-9 and np.abs(pgl[g:-g, g:-g] - pr[g:-g, g:-g]).max() <= 1e-9 * np.abs(pr).max()
    assert np.abs(vg[g:g + N1, g:-g, :] - vr[g:g + N1, g:-g, :]).max() <= 1e-8 * max(np.abs(vr).max(), 1e-30)
    # CalculateIntercellFlux_F / _G on slabs: this rank's rows of the whole-grid flux arrays against the 1-rank arrays
    ctx.fill_boundaries(); ref.fill_boundaries()
    for d in (0, 1):
        Fs = ctx.intercell_flux(d, dt_ref); Fr = ref.intercell_flux(d, dt_ref)
        own = slice(r0, r0 + n + 1) if d == 0 else slice(r0, r0 + n + 2 * g)
        assert np.abs(Fs[own] - Fr[own]).max() <= 1e-9 * max(np.abs(Fr).max(), 1e-30), (d, np.abs(Fs[own] - Fr[own]).max())
    if stype == O.PARALLEL:
        # parallel -> diagonal projection of a slab-decomposed field (collective; rows of the diagonal grid in chunks of the scratch field)
        for field, full in ((0, cr), (1, pr)):
            a = ctx.project_to_diagonal(field); b = ref.project_to_diagonal(field)
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-9 * max(np.abs(full).max(), 1.0), (field, np.abs(a - b).max())
    if os.environ.get("IMPEC_TEST_CKPT"):
        # per-rank binary checkpoint: save, step, restore, step again -> the same slab bit for bit
        import tempfile
        path = os.path.join(tempfile.gettempdir(), "impec_ckpt_%d_rank%d.bin" % (os.getppid(), rank))
        ctx.checkpoint_save(path, t, steps)
        dt1, _ = ctx.impes_step(t, 0.0)
        c1 = np.zeros_like(c); ctx.download_c(c1)
        assert ctx.checkpoint_load(path) == (t, steps)
        dt2, _ = ctx.impes_step(t, 0.0)
        c2 = np.zeros_like(c); ctx.download_c(c2)
        os.remove(path)
        assert dt1 == dt2 and np.array_equal(c1[rows], c2[rows]), "restart from the per-rank checkpoint is not bit-exact"
    if os.environ.get("IMPEC_REQUIRE_OVERLAP"):
        assert ctx.overlapped_exchanges() > 0, "no halo exchange was overlapped with interior chunks"
    print("rank %d ok: dc %.2e dp %.2e its %d/%d levels %d" % (rank, dc, dp, info.iterations, info_ref.iterations, ctx.mg_levels()), flush=True)
    ctx.close(); ref.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
