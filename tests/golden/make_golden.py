"""Regenerates tests/golden/oracle_vectors.npz: outputs of the CPU oracle (oracle/, the pinned restatement of the reference
path) on small seeded inputs.  The reference ships no golden files and cannot be built here, so these
vectors freeze the ORACLE: tests/test_golden.py checks that neither the oracle nor the kernel library drifts from them.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import oracle as O                                     # noqa: E402
from helpers import production_case, advection_case   # noqa: E402


def vectors():
    lib = O.lib()
    out = {}
    # one full IMPES step of the parallel.cfg operator at 24^2 (a20)
    base, Po, c, q = production_case(24, 24, stype=O.PARALLEL, eps=1e-12)
    S = O.new_state(); p = O.scalar_field(Po); u = O.velocity_field(Po); v = O.velocity_field(Po); co = c.copy()
    dt = lib.orc_impes_step(C.byref(Po), C.byref(S), p, co, u, v, q, 0.0, 0.0)
    out.update(impes_c0=c, impes_q=q, impes_c1=co, impes_p=p, impes_u=u, impes_v=v, impes_dt=np.array([dt]), impes_its=np.array([S.last_its]))
    # one MUSCL symmetric-parallel transport step with a perturbed velocity at 20^2 (a14-a17), bit-exact path
    base, Pa, ca, ua, va = advection_case(20)
    rng = np.random.default_rng(7)
    ua += 0.4 * rng.random(ua.shape); va -= 0.3 * rng.random(va.shape)
    dta = lib.orc_max_stable_dt(C.byref(Pa), ua, va)
    Sa = O.new_state(); cao = ca.copy()
    lib.orc_transport_step(C.byref(Pa), C.byref(Sa), cao, ua, va, np.zeros(400), dta)
    out.update(adv_c0=ca, adv_u=ua, adv_v=va, adv_c1=cao, adv_dt=np.array([dta]))
    # parallel -> diagonal projection of the synthetic state (8f rank 3)
    n1 = C.c_int(); n2 = C.c_int()
    lib.orc_project_dims(C.byref(Po), C.byref(n1), C.byref(n2))
    cf = c.copy(); lib.orc_fill_boundaries(C.byref(Po), cf)
    proj = np.zeros((n1.value, n2.value)); lib.orc_project_parallel(C.byref(Po), cf, n1.value, n2.value, proj)
    out.update(proj_in=cf, proj_out=proj)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **vectors())
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"))
