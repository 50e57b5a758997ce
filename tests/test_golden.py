"""Golden vectors (tests/golden/oracle_vectors.npz, written by tests/golden/make_golden.py from the pinned CPU oracle):
the oracle must still reproduce them bit for bit, and the kernel library (CPU-emulated here, the real one under -m gpu) must match
them to the parity tolerances WITHOUT the oracle in the loop."""
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import impec2d_b200 as I
import oracle as O
from helpers import production_case, advection_case, aos, interior

G = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))


def test_oracle_reproduces_golden_vectors():
    import make_golden
    now = make_golden.vectors()
    assert set(now) == set(G.files)
    for k in G.files:
        if k == "impes_its":                      # the PCG stops at a threshold: OpenMP reduction order can move it by one iteration
            assert abs(int(now[k][0]) - int(G[k][0])) <= 1
        elif k in ("impes_c1", "impes_p", "impes_u", "impes_v", "impes_dt"):          # behind the solve (eps = 1e-12): solver tolerance
            assert np.abs(now[k] - G[k]).max() <= 1e-9 * max(1.0, np.abs(G[k]).max()), k
        else:                                     # everything explicit: bit for bit
            assert np.array_equal(now[k], G[k]), k


def library_vs_golden(emulated):
    base, Po, c, q = production_case(24, 24, stype=O.PARALLEL, eps=1e-12)
    ctx = I.Context(I.default_params(emulated=emulated, **base), emulated=emulated)
    ctx.upload_c(G["impes_c0"]); ctx.upload_q(G["impes_q"])
    dt, info = ctx.impes_step(0.0, 0.0)
    assert abs(dt - G["impes_dt"][0]) <= 1e-9 * G["impes_dt"][0] and abs(info.iterations - int(G["impes_its"][0])) <= 1
    pg = ctx.download_p(); pr = G["impes_p"]
    assert np.linalg.norm(interior(Po, pg) - interior(Po, pr)) <= 1e-9 * np.linalg.norm(interior(Po, pr))
    assert np.abs(ctx.download_c() - G["impes_c1"]).max() <= 1e-8
    ctx.upload_c(G["proj_in"])
    assert np.array_equal(ctx.project_to_diagonal(0), G["proj_out"])
    ctx.close()
    base, Pa, ca, ua, va = advection_case(20)
    ctx = I.Context(I.default_params(emulated=emulated, **base), emulated=emulated)
    ctx.upload_c(G["adv_c0"]); ctx.upload_q(np.zeros(400)); ctx.upload_velocity(aos(G["adv_u"], G["adv_v"]))
    assert ctx.max_stable_dt() == G["adv_dt"][0]
    ctx.transport_step(float(G["adv_dt"][0]))
    assert np.array_equal(ctx.download_c(), G["adv_c1"])          # explicit transport path: bit-exact
    ctx.close()


def test_emulated_library_matches_golden_vectors(): library_vs_golden(True)


@pytest.mark.gpu
def test_gpu_library_matches_golden_vectors(): library_vs_golden(False)
