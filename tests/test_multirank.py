"""Slab decomposition across ranks on CPU: world_size-2 and -4 gloo runs of the emulated kernel library.
Each rank owns N1/P rows; halo rows travel by send/recv, norms by all-reduce, the first replicated multigrid level by
all-gather.  The slab run must reproduce the single-rank run (same kernels) to 1e-9."""
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_pool = None
_futures = {}
_calls = []          # the arguments of every run(...) call of this module's tests (filled by _collect)
_collecting = False


def _collect():
    """Dry pass over the tests of this module (with their parametrizations): run() only records its arguments."""
    global _collecting
    _collecting = True
    try:
        for name, fn in list(globals().items()):
            if not name.startswith("test_") or not callable(fn):
                continue
            marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
            if not marks:
                fn()
                continue
            names = [n.strip() for n in marks[0].args[0].split(",")]
            for vals in marks[0].args[1]:
                fn(**dict(zip(names, vals if isinstance(vals, (tuple, list)) else (vals,))))
    finally:
        _collecting = False


def _launch(world, N1, N2, stype, diffusion, steps, port, extra_env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py"), str(N1), str(N2), str(stype), str(diffusion), str(steps)]
    env = dict(os.environ, OMP_NUM_THREADS="1", IMPEC_DIST_MIN_CELLS="0")
    env.update(extra_env or {})
    for k in [k for k, v in env.items() if v is None]:
        del env[k]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)


def run(world, N1, N2, stype, diffusion, steps, port, extra_env=None):
    """One scenario = one torchrun launch of tests/multirank_worker.py (distinct rendezvous port each).  Most of a launch is start-up
    (three or five interpreters importing torch), so the first call starts ALL scenarios of this module on a small thread pool and every
    test waits for its own (8 cores: 105 s for the module instead of 290 s one after the other)."""
    global _pool
    if _collecting:
        _calls.append((world, N1, N2, stype, diffusion, steps, port, extra_env))
        return
    key = (world, N1, N2, stype, diffusion, steps, port, tuple(sorted((extra_env or {}).items(), key=lambda kv: kv[0])))
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "impec2d_b200", "csrc"), "emu"])
        _pool = ThreadPoolExecutor(max_workers=max(1, min(6, (os.cpu_count() or 2) // 2)))
        _collect()
        for args in _calls:
            k = args[:7] + (tuple(sorted((args[7] or {}).items(), key=lambda kv: kv[0])),)
            _futures[k] = _pool.submit(_launch, *args)
    if key not in _futures:                                  # a scenario that is not in CASES: run it now
        _futures[key] = _pool.submit(_launch, world, N1, N2, stype, diffusion, steps, port, extra_env)
    out = _futures[key].result()
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" ok: ") == world, out.stdout


def test_two_ranks_parallel_cfg(): run(2, 64, 64, 0, 0, 1, 29611)
def test_two_ranks_rect_diagonal_with_dispersion(): run(2, 64, 32, 1, 1, 1, 29612, {"IMPEC_TEST_CKPT": "1"})   # + per-rank checkpoint round trip
# the marching kernels, including the one-march Crank-Nicolson preconditioner (k_sgs_march), forced onto the slabs of a dispersion run
def test_two_ranks_dispersion_marching_kernels(): run(2, 64, 64, 1, 1, 1, 29615, {"IMPEC_MARCH_MIN_CELLS": "0"})
def test_four_ranks(): run(4, 64, 32, 1, 0, 1, 29613)
# the register-march multigrid kernels (big levels only by default) forced onto the slab levels: chunk edges at slab
# boundaries, the 2 halo rows PRE stores for its neighbours, coarse slab views at the transition level
def test_two_ranks_marching_kernels(): run(2, 128, 128, 0, 0, 1, 29614, {"IMPEC_MARCH_MIN_CELLS": "0"})


def test_two_ranks_default_replication_threshold():
    """WITHOUT IMPEC_DIST_MIN_CELLS: on a small grid every coarse level falls under the default 4M-cell threshold;
    level 0 must still be distributed (its vectors alias the slab-sized PCG vectors) and everything below replicated."""
    run(2, 64, 64, 0, 0, 1, 29616, {"IMPEC_DIST_MIN_CELLS": None})


def test_two_ranks_overlapped_halo_exchange():
    """Halo exchange overlapped with interior compute: with 8-row chunks a 64-row slab has interior chunks, so the marching
    PRE / POST / PLAIN / mat-vec kernels run as exchange + edge chunks + interior chunks; the result must equal the 1-rank run.  The same
    split with the exchange switched back to the serial order (IMPEC_HALO_OVERLAP=0) is covered by test_two_ranks_marching_kernels."""
    run(2, 128, 128, 0, 0, 1, 29617, {"IMPEC_MARCH_MIN_CELLS": "0", "IMPEC_MARCH_ROWS": "8", "IMPEC_REQUIRE_OVERLAP": "1"})


def test_two_ranks_horizontal_front():
    """source_type "horizontal": the Dirichlet row i = 1 lives on rank 0, the Crank-Nicolson Dirichlet row i = N1-2 on the last rank; the
    coarse-level shifts of the pinned row are computed on rank 0 and broadcast."""
    run(2, 64, 48, 2, 1, 1, 29618)


@pytest.mark.parametrize("split,port", [(0, 29619), (3, 29620), (2, 29621)])
def test_two_ranks_sequential_splittings(split, port):
    """dimensional_splitting = alternating_sweeps / strang / y_sweep_first on slabs: an x sweep is followed by an
    exchange of the swept rows, a y sweep recomputes the halo rows; two steps so that the alternating mode runs both orders."""
    run(2, 64, 64, 1, 0, 2, port, {"IMPEC_TEST_SPLITTING": str(split)})


def test_two_ranks_upwind_x_first():
    run(2, 64, 32, 1, 0, 2, 29622, {"IMPEC_TEST_SPLITTING": "1", "IMPEC_TEST_UPWIND": "1"})


def test_two_ranks_anisotropic_dispersion():
    """alpha_l != alpha_t on slabs: the matrix-free non-symmetric 9-point Crank-Nicolson operator reads one exchanged halo row of its
    operand per pass and folds ghost entries into the diagonal only at physical walls."""
    run(2, 64, 48, 1, 1, 2, 29623, {"IMPEC_TEST_ALPHA_T": "0.1"})


def test_four_ranks_parallel_projection_in_chunks():
    """parallel.cfg on four slabs; the 45 x 45 diagonal grid of Impec2DSolver::Project does not fit the 16-row slab's scratch field and is
    projected in two chunks of rows, each summed over the ranks (worker: projection of c and p against the 1-rank projection)."""
    run(4, 64, 64, 0, 0, 1, 29624)


@pytest.mark.parametrize("N1,N2,testno,nu,port", [(32, 48, 4, "1.0", 29625), (32, 48, 2, "0.6666666666666666", 29626), (20, 30, 4, "1.0", 29627)])
def test_two_ranks_elliptic_self_test(N1, N2, testno, nu, port):
    """The manufactured elliptic test (elliptic_test.cfg) on two slabs: the K(x,y) test fields are evaluated at global coordinates
    (the device-side x origin is the slab's), the multi-pass operator build exchanges its link rows."""
    run(2, N1, N2, 1, 0, 1, port, {"IMPEC_TEST_ELLIPTIC": str(testno), "IMPEC_TEST_NU": nu})


@pytest.mark.parametrize("world,N,method,port", [(2, 48, "muscl", 29628), (4, 64, "muscl", 29629), (2, 32, "upwind", 29630)])
def test_periodic_self_test_on_slabs(world, N, method, port):
    """hyperbolic_test.cfg (periodic advection) on 2 and 4 slabs: ghost rows from the neighbouring rank with the ranks closed to a ring
    (two ranks: both neighbours are the same peer, messages told apart by direction); bit-identical to the 1-rank run."""
    run(world, N, N, 1, 0, 5, port, {"IMPEC_TEST_PERIODIC": method})
