"""Multi-GPU parity (-m gpu, needs >= 2 devices): the slab-decomposed run over NCCL must reproduce the single-GPU run."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch
    return torch.cuda.device_count()


# march = 1: the register-march multigrid kernels (by default only on levels of >= 8M cells) forced onto every slab level
@pytest.mark.parametrize("N1,N2,stype,diffusion,steps,march", [(512, 512, 0, 0, 2, 0), (256, 128, 1, 1, 2, 0), (2048, 2048, 0, 0, 1, 0), (2048, 2048, 0, 0, 1, 1),
                                                                 (256, 192, 2, 1, 2, 0)])      # (the last: stype 2 = horizontal front, Dirichlet rows on the first / last rank)
def test_slabs_match_single_gpu(N1, N2, stype, diffusion, steps, march, split=None, extra_env=None):
    world = min(ngpus(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    while N1 % (2 * world) or N1 // world < 8:
        world //= 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tests", "multirank_worker.py"), str(N1), str(N2), str(stype), str(diffusion), str(steps), "cuda"]
    env = dict(os.environ, IMPEC_DIST_MIN_CELLS="0" if N1 < 2048 else "1000000")
    if march:
        env["IMPEC_MARCH_MIN_CELLS"] = "0"
    if split is not None:
        env["IMPEC_TEST_SPLITTING"] = str(split)
    env.update(extra_env or {})
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" ok: ") == world, out.stdout


@pytest.mark.parametrize("split", [0, 3])
def test_slabs_sequential_splittings(split):
    """alternating_sweeps / strang on slabs (added after the last GPU session of round 2: run on CPU ranks only so far,
    tests/test_multirank.py::test_two_ranks_sequential_splittings; the exchange after an x sweep is the same exchange_rows as everywhere)."""
    test_slabs_match_single_gpu(256, 128, 1, 0, 2, 0, split=split)


def test_slabs_self_tests():
    """The two self tests on slabs over NCCL (added after the last GPU session of round 2; CPU twins: test_two_ranks_elliptic_self_test,
    test_periodic_self_test_on_slabs): the manufactured elliptic test against the one-GPU solve, the periodic advection test (ranks
    closed to a ring, plain NCCL send / recv) bit for bit against the one-GPU run."""
    test_slabs_match_single_gpu(256, 384, 1, 0, 1, 0, extra_env={"IMPEC_TEST_ELLIPTIC": "4"})
    test_slabs_match_single_gpu(256, 256, 1, 0, 5, 0, extra_env={"IMPEC_TEST_PERIODIC": "muscl"})


@pytest.mark.parametrize("cfg_name,n1,n2", [("diagonal_dispersion", 256, 128), ("parallel", 256, 256)])
def test_cpp_driver_on_two_gpus(tmp_path, cfg_name, n1, n2):
    """The reference-facing C++ driver (`impec2d cfg`, Impec2DSolver(cfg, myid, num_procs)) as two processes on two GPUs (ranks from the
    environment, NCCL id through a file) against the same driver on one GPU: diagnostics and the final concentration / pressure /
    velocity files (gathered from both slabs by rank 0) agree to 1e-9.  parallel.cfg: c and p are written PROJECTED onto the diagonal grid
    (Impec2DSolver::Project, collective over the slabs) exactly as the one-GPU run writes them."""
    import re
    import numpy as np
    if ngpus() < 2:
        pytest.skip("needs at least 2 GPUs")
    lib_dir = os.path.join(ROOT, "impec2d_b200")
    subprocess.check_call(["make", "-s", "-C", os.path.join(lib_dir, "host"), "all"])
    exe = os.path.join(lib_dir, "impec2d")
    text = open(os.path.join(ROOT, "examples", cfg_name + ".cfg")).read()
    text = re.sub(r"x_cells\s*=\s*\d+", "x_cells = %d" % n1, text); text = re.sub(r"y_cells\s*=\s*\d+", "y_cells = %d" % n2, text)
    outs = {}
    for world in (1, 2):
        d = tmp_path / ("w%d" % world); d.mkdir()
        cfg = d / "run.cfg"; cfg.write_text(text.replace('"/tmp', '"' + str(d)))
        procs = []
        for r in range(world):
            env = dict(os.environ, IMPEC_RANK=str(r), IMPEC_NRANKS=str(world), IMPEC_DEVICE=str(r), IMPEC_NCCL_ID_FILE=str(d / "nccl_id"), IMPEC_DIST_MIN_CELLS="0")
            procs.append(subprocess.Popen([exe, str(cfg), "--steps", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
        logs = [p.communicate(timeout=600)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(l[-2000:] for l in logs)
        assert "[3]\tITS_e:" in logs[0], logs[0][-2000:]
        res = [os.path.join(dp, f) for dp, _, fs in os.walk(str(d)) for f in fs if f.endswith(".dat")]
        outs[world] = {os.path.basename(f): np.loadtxt(f) for f in res}
    assert {"concentration.dat", "pressure.dat", "velocity_u.dat", "velocity_v.dat", "diagnostics.dat"} <= set(outs[2]), sorted(outs[2])
    for name, a in outs[1].items():
        b = outs[2][name]
        assert a.shape == b.shape, name
        assert np.abs(a - b).max() <= 1e-9 * max(np.abs(a).max(), 1.0), (name, np.abs(a - b).max())
