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
@pytest.mark.parametrize("N1,N2,stype,diffusion,steps,march", [(512, 512, 0, 0, 2, 0), (256, 128, 1, 1, 2, 0), (2048, 2048, 0, 0, 1, 0), (2048, 2048, 0, 0, 1, 1)])
def test_slabs_match_single_gpu(N1, N2, stype, diffusion, steps, march):
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
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(" ok: ") == world, out.stdout
