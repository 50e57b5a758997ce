#!/usr/bin/env python
"""Race check without a GPU: the emulated kernels run with blocks / threads / coroutines in forward, in REVERSE and in a pseudo-random order
(IMPEC_EMU_ORDER, csrc/emu.cpp) must produce the same BITS.  A kernel that reads what another thread or block of the same launch
writes (in-place stencil, missing barrier) would differ.  Scenarios: IMPES steps with and without dispersion on awkward extents, the
anisotropic-dispersion path, the sequential splittings with upwind fluxes, the horizontal front, marching kernels forced onto
every level, the periodic self test.  Usage: python tools/emu_order_check.py"""
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    dict(name="impes_parallel_96", N1=96, N2=96, stype=0, diffusion=0, steps=2),
    dict(name="impes_diag_disp_100x50", N1=100, N2=50, stype=1, diffusion=1, steps=2),
    dict(name="impes_diag_aniso_64x48", N1=64, N2=48, stype=1, diffusion=1, steps=1, alpha_t=0.1),
    dict(name="impes_strang_upwind_72x40", N1=72, N2=40, stype=1, diffusion=0, steps=2, dim_splitting=3, hyper_method=1),
    dict(name="impes_horizontal_64x48", N1=64, N2=48, stype=2, diffusion=1, steps=2),
    dict(name="impes_march_everywhere_128", N1=128, N2=128, stype=0, diffusion=1, steps=1, env={"IMPEC_MARCH_MIN_CELLS": "0", "IMPEC_MARCH_ROWS": "8"}),
    dict(name="periodic_50", N1=50, N2=50, periodic=True, steps=6),
]


def child(case_idx, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import impec2d_b200 as I
    from helpers import production_case, advection_case, aos
    cs = CASES[case_idx]
    if cs.get("periodic"):
        base, Po, c, u, v = advection_case(cs["N1"])
        ctx = I.Context(I.default_params(emulated=True, **base), emulated=True)
        ctx.upload_c(c); ctx.upload_velocity(aos(u, v)); ctx.upload_q(np.zeros(cs["N1"] * cs["N2"]))
        dt = ctx.max_stable_dt()
        for _ in range(cs["steps"]):
            ctx.transport_step(dt)
        np.savez(out, c=ctx.download_c(), dt=dt)
        return
    kw = {k: cs[k] for k in ("alpha_t", "dim_splitting", "hyper_method") if k in cs}
    base, Po, c, q = production_case(cs["N1"], cs["N2"], stype=cs["stype"], diffusion=cs["diffusion"], eps=1e-12, **kw)
    ctx = I.Context(I.default_params(emulated=True, **base), emulated=True)
    ctx.upload_c(c); ctx.upload_q(q)
    t, its = 0.0, []
    for _ in range(cs["steps"]):
        dt, info = ctx.impes_step(t, 0.0); t += dt; its.append(info.iterations)
    d = ctx.diagnostics()
    np.savez(out, c=ctx.download_c(), p=ctx.download_p(), v=ctx.download_velocity(), t=t, its=np.array(its), mass=d.total_mass, outflux=d.total_mass_outflux)


def main():
    if len(sys.argv) == 4 and sys.argv[1] == "--child":
        return child(int(sys.argv[2]), sys.argv[3])
    import tempfile
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "impec2d_b200", "csrc"), "emu"])
    bad = 0
    with tempfile.TemporaryDirectory() as d:
        for k, cs in enumerate(CASES):
            outs = []
            for order in ("forward", "reverse", "shuffle"):
                out = os.path.join(d, "%s_%s.npz" % (cs["name"], order))
                env = dict(os.environ, IMPEC_EMU_ORDER=order, OMP_NUM_THREADS="1"); env.update(cs.get("env", {}))
                subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", str(k), out], env=env)
                outs.append(np.load(out))
            same = all(np.array_equal(outs[0][f], o[f]) for f in outs[0].files for o in outs[1:])
            worst = max(float(np.abs(np.asarray(outs[0][f], dtype=float) - np.asarray(o[f], dtype=float)).max()) for f in outs[0].files for o in outs[1:])
            print("%-34s %s (max difference %.3e; fields %s)" % (cs["name"], "bit-identical" if same else "DIFFERS", worst, ",".join(outs[0].files)))
            bad += not same
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
