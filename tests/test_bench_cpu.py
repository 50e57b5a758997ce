"""bench.py without a GPU: the reference arm (`--impl reference`, the CPU implementations timed on the host cores) on small grids, the
behaviour of the arm under torchrun's environment (rank > 0 prints nothing, OMP_NUM_THREADS = 1 is ignored), the byte model helpers, and
that our own arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "cpu_baseline", "e2e")


def run_bench(*args, env=None):
    out = subprocess.run([sys.executable, BENCH] + list(args), capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    return out


def test_reference_arm_transport_line():
    """hyperbolic_test through the reference's own code where oracle/_ref is built, else the oracle port; one JSON line, contract keys."""
    out = run_bench("--impl", "reference", "--workload", "hyperbolic_test", "--size", "256", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    for k in REQUIRED:
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "Gcell-updates/s" and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 2 and line["warmup"] == 3       # W >= 3 is enforced
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["grid"] == [256, 256] and "hyperbolic_test.cfg" in line["config"]["workload"]
    assert line["gpu_launches"] == 0


def test_reference_arm_under_torchrun_environment():
    """torchrun exports OMP_NUM_THREADS=1: rank 0 still uses every host thread; the other ranks exit 0 without output."""
    env = {"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}
    out = run_bench("--impl", "reference", "--gpus", "2", "--workload", "hyperbolic_test", "--size", "128", "--steps", "1", "--warmup", "1", env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    out = run_bench("--impl", "reference", "--gpus", "2", "--workload", "hyperbolic_test", "--size", "128", "--steps", "1", "--warmup", "1",
                    env=dict(env, RANK="1", LOCAL_RANK="1"))
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_impes_sample():
    """The IMPES workloads run the oracle port (the reference's solve is HYPRE) on a bounded sample: at --size 2048 that is the grid itself."""
    out = run_bench("--impl", "reference", "--workload", "diagonal_dispersion", "--size", "2048", "--steps", "1", "--warmup", "0", "--eps", "1e-8")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    b = line["cpu_baseline"]
    assert b["kind"] == "port" and b["sample_grid"] == [2048, 2048] and len(b["pcg_iterations"]) == 1 and b["pcg_iterations"][0] > 0
    assert line["config"]["eps"] == 1e-8 and "diagonal_dispersion.cfg" in line["config"]["workload"]


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = run_bench("--steps", "1", "--warmup", "1", "--size", "256")
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_cycle_byte_model():
    """Bytes per fine cell of one multigrid cycle as bench.py derives them from the W range in effect."""
    sys.path.insert(0, ROOT)
    import bench
    assert abs(bench.cycle_bytes_per_fine_cell(2, 4, 13) - 113.3) < 0.05          # default on 1-4 ranks
    assert abs(bench.cycle_bytes_per_fine_cell(2, 3, 13) - 111.6) < 0.05          # default on >= 8 ranks
    v = bench.cycle_bytes_per_fine_cell(99, 99, 13)                               # plain V cycle: 76 * 4/3
    assert abs(v - 76.0 * (1 - 0.25 ** 13) / 0.75) < 1e-9
    for wl, w in bench.WORKLOADS.items():
        assert os.path.exists(os.path.join(ROOT, w["cfg"])), wl


@pytest.mark.parametrize("workload,size,extra", [("parallel", 128, []), ("diagonal_dispersion", 96, ["--no-cpu-baseline"]), ("hyperbolic_test", 128, ["--no-cpu-baseline"]),
                                                 ("parallel", 96, ["--profile-mode"])])
def test_our_arm_dry_run_on_the_emulator(workload, size, extra):
    """bench.py's own arm end to end -- warm-up, timed steps, per-kernel timing calls, byte models, e2e leg with host buffers, cpu_baseline leg,
    the JSON line -- with torch.cuda's streams / events stubbed and the CPU-emulated kernel library in place of the product one
    (tests/bench_dry_run.py).  Only the Python of the measurement path is under test; the numbers mean nothing."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dry_run.py"), "--workload", workload, "--size", str(size), "--steps", "1",
                          "--warmup", "1", "--eps", "1e-8"] + extra, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    if "--profile-mode" in extra:
        assert line["profile_mode"] is True and line["gpu_launches"] > 0
        return
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "run", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert "impl" not in line and line["n_gpus"] == 1 and line["warmup"] == 3 and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["gpu_launches"] > 0 and line["value"] > 0
    assert line["config"]["grid"] == [size, size] and "model" not in line["config"]
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0 and e["unit"] == line["unit"]
    if workload == "hyperbolic_test":
        assert line["metric"] == "MUSCL transport Gcell-updates/s" and "vcycle_ms" not in line
    else:
        assert line["metric"] == "IMPES Gcell-updates/s" and line["vcycle_ms"] > 0
        assert len(line["run"]["pcg_iterations"]) == 1 and line["run"]["mg_w_range"] == [2, 4]
        assert r["step"]["moved_bytes"] > 0 and r["vcycle"]["moved_bytes"] > 0
    if "--no-cpu-baseline" in extra:
        assert line["cpu_baseline"] is None
    else:
        b = line["cpu_baseline"]
        assert b["kind"] == "port" and b["cores"] >= 1 and b["value"] > 0 and b["sample_grid"] == [size, size]


def test_our_arm_dry_run_two_ranks(tmp_path):
    """The N > 1 path of bench.py (slab inputs, id broadcast, max / sum over ranks, host binding report, one line from rank 0) as two torchrun
    processes: gloo for torch.distributed, the emulated build's file communicator for the library, a made-up local CPU list for the binding."""
    cpus = sorted(os.sched_getaffinity(0))
    env = dict(os.environ, IMPEC_EMU_COMM_DIR=str(tmp_path), IMPEC_DIST_MIN_CELLS="0", IMPEC_EMU_LOCAL_CPULIST=",".join(str(c) for c in cpus[:2]))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29678",
           os.path.join(ROOT, "tests", "bench_dry_run.py"), "--gpus", "2", "--workload", "diagonal_dispersion", "--size", "96", "--steps", "1", "--warmup", "1", "--eps", "1e-8"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout                               # rank 0 alone prints
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["cpu_baseline"] is None
    assert line["run"]["host_binding_rank0"] == {"cpus": min(2, len(cpus)), "cpulist": env["IMPEC_EMU_LOCAL_CPULIST"]}
    assert "2 slabs of 48 rows" in line["run"]["parallelism"]
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0 and line["vcycle_ms"] > 0
