#!/usr/bin/env python
"""Random-shape parity fuzz of the kernel source on the CPU emulator against the oracle (no GPU needed):

    python tools/emu_fuzz.py SEED COUNT

draws COUNT cases -- an IMPES step (parallel / diagonal grid, any nu, either averaging), transport steps (MUSCL / upwind, any of the five
splittings, porosity 1 or not), a dispersion step (alpha_l = or != alpha_t) or a horizontal-front step -- on grids of 4..90 cells per
direction, odd, even and prime extents alike, and runs the tests' own checks (tests/parity_cases.py: bit-exact transport, pressure to
1e-9, iteration counts +-1) on each.  Exit code = number of failing cases.  Last run: seeds 1, 2, 3 x 40 and 4, 5, 6 x 30 cases: every field within tolerance; one case (24 x 64 cells, nu = 1/4) needed 82 PCG iterations
against the oracle's 64-66 with the FP32 coefficient copies (64 with FP64 coefficients) -- see check_impes_step."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O               # noqa: E402
import parity_cases as pc        # noqa: E402


def draw(rng):
    """(description, thunk) of one random case"""
    kind = int(rng.integers(0, 4))
    if kind == 0:
        stype = int(rng.choice([O.PARALLEL, O.DIAGONAL]))
        N1 = int(rng.integers(4, 90)); N2 = N1 if stype == O.PARALLEL else int(rng.integers(4, 90))        # parallel grids are square (src/SimParams.cpp:460-463)
        nu = float(rng.choice([1.0, 2 / 3, 0.5, 0.25])); avg = int(rng.choice([O.HARMONIC, O.MEANCONC]))
        hard = nu < 0.5 and max(N1, N2) > 2 * min(N1, N2)          # weakly preconditioned: iteration counts are not compared (parity_cases.check_impes_step)
        return ("impes", N1, N2, stype, nu, avg), lambda: pc.check_impes_step(True, N1, N2, stype=stype, nu=nu, avg=avg, steps=1, its_tol=None if hard else 1)
    if kind == 1:
        N = int(rng.integers(6, 70)); method = int(rng.choice([O.MUSCL, O.UPWIND])); split = int(rng.integers(0, 5)); por = float(rng.choice([1.0, 0.37]))
        return ("advection", N, method, split, por), lambda: pc.check_advection_bitexact(True, N=N, steps=3, method=method, split=split, porosity=por)
    if kind == 2:
        N1, N2 = int(rng.integers(6, 80)), int(rng.integers(6, 80)); at = float(rng.choice([1.0, 0.3]))
        return ("dispersion", N1, N2, at), lambda: pc.check_dispersion(True, N1, N2, stype=O.DIAGONAL, steps=1, alpha_l=1.0, alpha_t=at)
    N1, N2 = int(rng.integers(20, 80)), int(rng.integers(6, 70)); d = int(rng.integers(0, 2))        # (an injection layer of 0.05 narrower than one cell: the reference exits)
    return ("horizontal", N1, N2, d), lambda: pc.check_horizontal_front(True, N1, N2, diffusion=d, steps=1)


def main(seed, count):
    rng = np.random.default_rng(seed)
    fails = 0
    for _ in range(count):
        desc, run = draw(rng)
        try:
            run()
            print("ok  ", desc, flush=True)
        except Exception as e:
            fails += 1
            print("FAIL", desc, type(e).__name__, str(e)[:300], flush=True)
            traceback.print_exc(limit=3)
    print("failures:", fails)
    return fails


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 10))
