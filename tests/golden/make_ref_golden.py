"""Regenerates tests/golden/ref_vectors.npz: outputs of the REFERENCE ITSELF (pw359/Impec2D, its own translation units compiled
unmodified into oracle/_ref/libimpec_ref.so and oracle/_ref/Impec2D_ref, see oracle/Makefile and oracle/ref_api.cpp) on small seeded
inputs.  Needs /root/reference, so it runs in the build container only; the vectors are committed and travel to the GPU box, where
neither the reference sources nor (necessarily) the compiled reference exist.

    python tests/golden/make_ref_golden.py

Contents (inputs are stored next to the outputs, so a consumer needs nothing but this file):
  ell_*    parallel.cfg operator on a 24 x 24 developed state: the assembled 9-entry matrix (HYPRE entry order), right-hand side,
           converged pressure, Darcy velocity, stable time step
  tr_*     one transport step (TR source, MUSCL symmetric-parallel advection) on that velocity field + produced mass
  up_*     the same step with UpwindSplitSolver
  cn_*     diagonal_dispersion.cfg Crank-Nicolson half step at 20 x 28 (dx != dy): matrix, right-hand side, solution
  run_*    4 IMPES steps of parallel.cfg at 32 x 32 from c = 0 (c, p after every step, dt, produced mass)
  adv_*    hyperbolic_test.cfg: 10 periodic MUSCL steps of the sine profile at 32 x 32, CFL 0.95
  exe_*    the reference's OWN main() on elliptic_test.cfg (testno 4, 20x30 doubled twice) and hyperbolic_test.cfg (sine, 25^2 and
           50^2): the norm files it writes
"""
import ctypes as C
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import oracle as O                 # noqa: E402   (only for the seeded synthetic INPUT state and the well model, both checked bit for
from oracle import ref as R        # noqa: E402    bit against the reference in tests/test_oracle_vs_ref.py)

SQRT2 = 1.41421356237


def synth(N1, N2, **kw):
    Po = O.default_params(N1=N1, N2=N2, **kw)
    c = O.scalar_field(Po)
    O.lib().orc_synth_state(C.byref(Po), 20260917, c)
    return c


def vectors():
    out = {}
    # ---- ell / tr / up: parallel.cfg at 24 x 24
    N = 24
    r = R.Ref(general=dict(x_cells=N, y_cells=N))
    c0 = synth(N, N, x_ub=SQRT2, y_ub=SQRT2, stype=O.PARALLEL)
    r.set("c", c0); r.L.ref_fill_boundaries(r.h); c0 = r.get("c")
    r.L.ref_init_source(r.h, 0.0); q = r.get("q")
    r.L.ref_pressure_solve(r.h)
    A, b, x, rec = r.last_system()
    r.L.ref_calc_velocities(r.h)
    u, v = r.get_uv()
    dt = r.L.ref_max_stable_dt(r.h)
    out.update(ell_c=c0, ell_q=q, ell_A=A, ell_b=b, ell_p=r.get("p"), ell_u=u, ell_v=v, ell_dt=np.array([dt]))
    r.L.ref_transport_step(r.h, dt, 0)
    out.update(tr_c1=r.get("c"), tr_outflux=np.array([r.L.ref_total_mass_outflux(r.h, 0)]))
    r.set("c", c0)
    r.L.ref_transport_step(r.h, dt, 1)
    out.update(up_c1=r.get("c"), up_outflux=np.array([r.L.ref_total_mass_outflux(r.h, 1)]))
    r.close()
    # ---- cn: diagonal_dispersion.cfg operator on a 20 x 28 grid of the unit square
    N1, N2 = 20, 28
    r = R.Ref(general=dict(x_cells=N1, y_cells=N2, x_ub=1.0, y_ub=1.0, source_type="diagonal"), hyperbolic=dict(diffusion=True))
    c0 = synth(N1, N2, x_ub=1.0, y_ub=1.0, stype=O.DIAGONAL)
    r.set("c", c0); r.L.ref_fill_boundaries(r.h); c0 = r.get("c")
    r.L.ref_init_source(r.h, 0.0)
    r.L.ref_pressure_solve(r.h); r.L.ref_calc_velocities(r.h)
    u, v = r.get_uv()
    dth = 0.5 * r.L.ref_max_stable_dt(r.h)
    r.L.ref_diffusion_split(r.h, dth)
    A, b, x, rec = r.last_system()
    out.update(cn_c=c0, cn_u=u, cn_v=v, cn_dt=np.array([dth]), cn_A=A, cn_b=b, cn_c1=r.get("c"))
    r.close()
    # ---- run: 4 IMPES steps from c = 0
    N = 32
    r = R.Ref(general=dict(x_cells=N, y_cells=N))
    r.L.ref_init_source(r.h, 0.0)
    t = 0.0; cs, ps, dts = [], [], []
    for _ in range(4):
        its = C.c_int()
        dt = r.L.ref_impes_step(r.h, t, 0.0, C.byref(its)); t += dt
        cs.append(r.get("c")); ps.append(r.get("p")); dts.append(dt)
    out.update(run_q=r.get("q"), run_c=np.array(cs), run_p=np.array(ps), run_dt=np.array(dts), run_outflux=np.array([r.L.ref_total_mass_outflux(r.h, 0)]))
    r.close()
    # ---- adv: periodic advection
    N = 32
    r = R.Ref(general=dict(x_cells=N, y_cells=N, x_ub=1.0, y_ub=1.0, source_type="diagonal", is_test=True), hyperbolic=dict(cfl=0.95),
              test=dict(component="hyperbolic", testno=1))
    g = r.g
    c0 = np.zeros(r.scalar_shape())
    dx = 1.0 / N
    for i in range(N):
        for j in range(N):
            c0[g + i, g + j] = r.L.ref_profile(r.h, 0, (i + 0.5) * dx, (j + 0.5) * dx)
    vel = np.zeros(r.velocity_shape()); vel[:, :, 0] = -1.0; vel[:, :, 1] = 1.0
    r.set("c", c0); r.set("v", vel)
    dt = r.L.ref_max_stable_dt(r.h)
    for _ in range(10):
        r.L.ref_transport_step(r.h, dt, 0)
    out.update(adv_c0=c0, adv_dt=np.array([dt]), adv_c10=r.get("c"))
    r.close()
    # ---- exe: the reference's own main()
    with tempfile.TemporaryDirectory() as d:
        files = {k: os.path.join(d, k + ".out") for k in ("num_result_file", "ex_result_file", "num_u_file", "num_v_file", "ex_u_file", "ex_v_file")}
        txt = R.cfg_text(general=dict(x_ub=6.283185307179586476, y_ub=6.283185307179586476, source_type="diagonal", x_cells=50, y_cells=50, is_test=True,
                                      results_dir=d, log_dir=d),
                         elliptic=dict(eps=1e-12, mu_res=1., mu_inj=1., average_procedure="mean_concentration"), hyperbolic=dict(nu=1., diffusion=True),
                         test=dict(component="elliptic", testno=4, initial_x_cells=20, initial_y_cells=30, profile="tophat", hyp_norm=2, levels=3,
                                   norm_file=os.path.join(d, "ell_norm.out"), **files))
        R.run_executable(txt, cwd=d)
        out["exe_elliptic_norms"] = np.loadtxt(os.path.join(d, "ell_norm.out"))
        txt = R.cfg_text(general=dict(endtime=1.0, x_ub=1.0, y_ub=1.0, source_type="diagonal", x_cells=80, y_cells=80, is_test=True, results_dir=d, log_dir=d),
                         elliptic=dict(mu_res=1., mu_inj=1.), hyperbolic=dict(cfl=0.95, nu=1.),
                         test=dict(component="hyperbolic", testno=1, initial_x_cells=25, initial_y_cells=25, profile="sine", hyp_norm=1, levels=2,
                                   norm_file=os.path.join(d, "hyp_norm.out"), **files))
        R.run_executable(txt, cwd=d)
        out["exe_hyperbolic_norms"] = np.loadtxt(os.path.join(d, "hyp_norm.out"))
    return out


if __name__ == "__main__":
    v = vectors()
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **v)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, a in v.items():
        print("  %-22s %s" % (k, a.shape))
