"""Iteration counts of the restated MG-PCG for different cycle shapes (oracle, CPU)."""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle as O
from helpers import production_case
lib = O.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variants = [(2, 4), (2, 3), (1, 2), (1, 3), (0, 0), (0, 1), (0, 2), (1, 1)]
for wf, wt in variants:
    for nc in (8,):
        base, P, c, q = production_case(N, N, eps=1e-14, mg_w_from=wf, mg_w_to=wt, mg_ncoarse=nc)
        S = O.new_state(); p = O.scalar_field(P)
        t = time.time()
        rc = lib.orc_solve_elliptic(C.byref(P), C.byref(S), c, q, p)
        print(N, 'W levels', wf, wt, 'ncoarse', nc, 'rc', rc, 'its', S.last_its, 'true', S.last_true_relres, round(time.time() - t, 1), 's', flush=True)
