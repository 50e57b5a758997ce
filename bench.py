#!/usr/bin/env python
"""bench.py -- hot-path benchmark of the B200-native Impec2D path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--size S] [--eps E] [--impl ours|reference]

Workloads (BASELINE.json `configs`; all FP64, synthetic inputs of the named shape):
  parallel (default, the headline: "IMPES Gcell-updates/s & MG V-cycle ms at 16384^2")
      examples/parallel.cfg operator (9-point blend nu = 2/3, harmonic mobility, mu_res/mu_inj = 41, quarter-circle wells, MUSCL
      symmetric-parallel transport, CFL 0.5) on an S x S grid (default 16384), started from the developed synthetic fingering state of
      SURVEY.md 8(d).  One step = one iteration of Impec2DSolver::Solve's loop (src/Impec2DSolver.cpp:343-389): FillBoundaries, MG-PCG
      pressure solve from x0 = 0 to the cfg tolerance (1e-14, HYPRE's recurrence-residual test), CalcVelocities, MaxStableTimestep,
      transport step.
  diagonal_dispersion (BASELINE config 4, default 16384^2) / parallel_dispersion (config 5, default 32768^2: needs 8 GPUs)
      the same step with the velocity-dependent dispersion tensor: two Crank-Nicolson half steps per transport step
      (examples/diagonal_dispersion.cfg, examples/parallel_dispersion.cfg; src/MGSolver.cpp:846-976).
  hyperbolic_test (BASELINE config 2, default 4096^2, one GPU)
      examples/hyperbolic_test.cfg: periodic MUSCL symmetric-parallel advection of the sine profile with the prescribed velocity
      (-1, +1), CFL 0.95.  One step = HyperbolicSplitSolver::Solve, as in the loop of Impec2DSolver::Advection (src/Impec2DSolver.cpp:929-944).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same metric through the C ABI with
pinned HOST buffers (c and q handed over and c read back every step).  `--impl reference` times the reference's CPU implementation
of the same path on all host threads: for hyperbolic_test the reference's OWN code (oracle/_ref: pw359/Impec2D's sources compiled
unmodified, its OpenMP switches on); for the IMPES workloads the C restatement (oracle/, OpenMP) -- the reference's own pressure
solve lives in HYPRE, which is not available -- on a bounded sample of the workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SQRT2 = 1.41421356237
CPU_BUDGET_REFERENCE_S = 180.0   # `--impl reference`: the sample is sized so that warm-up + timed steps stay within about this much CPU time
CPU_BUDGET_BASELINE_S = 30.0     # `cpu_baseline` inside our own arm's line: a bounded sample of 10-30 s of CPU work

WORKLOADS = {
    "parallel": dict(kind="impes", size=16384, L=SQRT2, stype=0, diffusion=0, cfg="examples/parallel.cfg"),
    "diagonal_dispersion": dict(kind="impes", size=16384, L=1.0, stype=1, diffusion=1, cfg="examples/diagonal_dispersion.cfg"),
    "parallel_dispersion": dict(kind="impes", size=32768, L=SQRT2, stype=0, diffusion=1, cfg="examples/parallel_dispersion.cfg"),
    "hyperbolic_test": dict(kind="transport", size=4096, L=1.0, stype=1, diffusion=0, cfg="examples/hyperbolic_test.cfg"),
}


def workload_params(wl, N, eps, module):
    w = WORKLOADS[wl]
    if w["kind"] == "transport":
        return module.default_params(N1=N, N2=N, x_ub=1.0, y_ub=1.0, stype=1, is_test=1, testno=1, cfl=0.95, diffusion=0, hyper_method=0,
                                     dim_splitting=4, mu_res=41.0, mu_inj=1.0, nu=2.0 / 3.0)
    return module.default_params(N1=N, N2=N, x_ub=w["L"], y_ub=w["L"], stype=w["stype"], nu=2.0 / 3.0, mu_res=41.0, mu_inj=1.0, avgproc=1, cfl=0.5,
                                 eps=eps, para_eps=1e-14, hyper_method=0, dim_splitting=4, diffusion=w["diffusion"], alpha_l=1.0, alpha_t=1.0, ptype=1)


def workload_text(wl, N):
    w = WORKLOADS[wl]
    if w["kind"] == "transport":
        return "%s transport step (periodic MUSCL symmetric-parallel advection, sine profile, velocity (-1,+1), CFL 0.95) at %d^2" % (w["cfg"], N)
    extra = " + 2 Crank-Nicolson dispersion half steps" if w["diffusion"] else ""
    return "%s IMPES step (MG-PCG pressure to eps + velocity + CFL + MUSCL transport%s) at %d^2, developed synthetic state" % (w["cfg"], extra, N)


def sine_cell_averages(N):
    """Sinewave of src/Util.cpp:94-98 at the cell centres of the unit square (a product of two 1-D factors)."""
    import numpy as np
    dx = 1.0 / N
    x = (np.arange(N) + 0.5) * dx
    f = np.cos(2.0 * np.pi * (x + dx / 2.0)) - np.cos(2.0 * np.pi * (x - dx / 2.0))
    return 1.0 + 1.0 / (dx * dx * 4.0 * np.pi * np.pi) * np.outer(f, f)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.samples = []; self.stop_flag = False; self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append(line.strip())
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All the host's hardware threads this process may use -- NOT OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# -------------------------------------------------------------------------------------------------------- CPU arms
def cpu_impes_run(wl, steps, warmup, eps, n):
    """The oracle (C restatement of the reference path, OpenMP over all host threads): IMPES steps on an n x n grid of the workload.
    Returns (Gcell-updates/s, ms per step, threads, PCG iterations per step)."""
    import numpy as np
    import oracle as O
    lib = O.lib()
    lib.orc_set_num_threads(host_threads())
    P = workload_params(wl, n, eps, O)
    c = O.scalar_field(P); lib.orc_synth_state(C.byref(P), 20260917, c)
    q = np.zeros(n * n); lib.orc_init_source(C.byref(P), 1.0, 0.05, 0, q)
    p = O.scalar_field(P); u = O.velocity_field(P); v = O.velocity_field(P)
    S = O.new_state()
    t = 0.0
    for _ in range(warmup):
        t += lib.orc_impes_step(C.byref(P), C.byref(S), p, c, u, v, q, t, 0.0)
    its = []
    t0 = time.perf_counter()
    for _ in range(steps):
        t += lib.orc_impes_step(C.byref(P), C.byref(S), p, c, u, v, q, t, 0.0)
        its.append(S.last_its)
    el = time.perf_counter() - t0
    return n * n * steps / el / 1e9, el / steps * 1e3, lib.orc_num_threads(), its


def cpu_impes_sample_size(wl, steps, warmup, eps, full, budget_s):
    """Largest of {full, 8192, 4096, 2048} whose warm-up + timed steps fit budget_s, from a calibration step at 2048^2."""
    val, ms, cores, its = cpu_impes_run(wl, 1, 0, eps, 2048)
    s_per_cell = ms * 1e-3 / (2048.0 * 2048.0)
    for n in (full, 8192, 4096, 2048):
        if n <= full and s_per_cell * n * n * (steps + warmup) * 1.15 <= budget_s:        # 1.15: bigger grids run a little slower per cell
            return n
    return 2048


def cpu_transport_run(steps, warmup, n):
    """hyperbolic_test.cfg transport steps on the CPU.  Preferred: the REFERENCE's own code (oracle/_ref/libimpec_ref_omp.so =
    pw359/Impec2D's MUSCLSplitSolver etc. compiled unmodified with its OpenMP switches on); otherwise the oracle port."""
    import numpy as np
    from oracle import ref as R
    c0 = np.zeros((n + 4, n + 4)); c0[2:-2, 2:-2] = sine_cell_averages(n)
    dt = 0.95 / n
    if R.available(omp=True):
        r = R.Ref(omp=True, general=dict(x_cells=n, y_cells=n, x_ub=1.0, y_ub=1.0, source_type="diagonal", is_test=True), hyperbolic=dict(cfl=0.95),
                  test=dict(component="hyperbolic", testno=1))
        r.L.ref_set_threads(host_threads())
        vel = np.zeros(r.velocity_shape()); vel[:, :, 0] = -1.0; vel[:, :, 1] = 1.0
        r.set("c", c0); r.set("v", vel)
        for _ in range(warmup):
            r.L.ref_transport_step(r.h, dt, 0)
        t0 = time.perf_counter()
        for _ in range(steps):
            r.L.ref_transport_step(r.h, dt, 0)
        el = time.perf_counter() - t0
        r.close()
        return n * n * steps / el / 1e9, el / steps * 1e3, host_threads(), "reference"
    import oracle as O
    lib = O.lib(); lib.orc_set_num_threads(host_threads())
    P = workload_params("hyperbolic_test", n, 0.0, O)
    u = O.velocity_field(P, -1.0); v = O.velocity_field(P, 1.0); q = np.zeros(n * n); S = O.new_state(); c = c0.copy()
    for _ in range(warmup):
        lib.orc_transport_step(C.byref(P), C.byref(S), c, u, v, q, dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.orc_transport_step(C.byref(P), C.byref(S), c, u, v, q, dt)
    el = time.perf_counter() - t0
    return n * n * steps / el / 1e9, el / steps * 1e3, lib.orc_num_threads(), "port"


def cpu_arm(wl, N, steps, warmup, eps, budget_s=CPU_BUDGET_REFERENCE_S):
    """(value, ms/step, cpu_baseline dict, extra config) of the CPU implementation of workload wl."""
    if WORKLOADS[wl]["kind"] == "transport":
        val, ms, cores, kind = cpu_transport_run(steps, warmup, N)
        what = ("the reference's own code (oracle/_ref: src/MUSCLSplitSolver.cpp, HyperbolicSplitSolver.cpp, TRIntegrator.cpp, Util.cpp compiled unmodified, "
                "-DOMPHYPER -DOMPUTIL, OpenMP over all host threads)") if kind == "reference" else "CPU restatement of the reference path (oracle/), OpenMP over all host threads"
        return val, ms, {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": kind, "ms_per_step": ms,
                         "sample": "%d transport steps at the FULL %d^2 grid of the workload after %d warm-up; %s" % (steps, N, warmup, what)}, {}
    n = cpu_impes_sample_size(wl, steps, warmup, eps, N, budget_s)
    val, ms, cores, its = cpu_impes_run(wl, steps, warmup, eps, n)
    return val, ms, {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port", "ms_per_step": ms, "pcg_iterations": its, "sample_grid": [n, n],
                     "sample": "%d IMPES steps on a %d^2 sample of the workload (same operator, eps and state generator; the largest of 16384/8192/4096/2048 "
                               "whose %d steps fit ~%d s of CPU time) after %d warm-up; C restatement of the reference path (oracle/; the reference's own solve "
                               "is HYPRE, unavailable), OpenMP over all %d host threads" % (steps, n, steps + warmup, int(budget_s), warmup, cores)}, {}


def common_config(wl, N, eps):
    """The part of `config` both arms share (the driver compares the two)."""
    cfg = {"workload": workload_text(wl, N), "grid": [N, N],
           "l2": "inputs larger than L2 (every field is %.2f GB)" % (8.0 * N * N / 1e9)}
    if WORKLOADS[wl]["kind"] == "impes":
        cfg.update({"eps": eps, "stop_rule": "<r,r>/<b,b> < eps^2, recurrence residual, x0 = 0 (HYPRE StructPCG semantics)"})
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="parallel", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="grid size (default: the workload's BASELINE size)")
    ap.add_argument("--eps", type=float, default=1e-14, help="pressure tolerance (examples/parallel.cfg:52)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="library option for A/B runs (impec_set_option), e.g. mg_w_to=3; recorded in the line's run.options")
    ap.add_argument("--profile-mode", action="store_true", help="for runs under ncu: honour --warmup as given, skip e2e / cpu baseline / kernel timing (numbers printed in this mode are not bench values)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = args.warmup if args.profile_mode else max(args.warmup, 3)
    if args.profile_mode:
        args.no_e2e = True; args.no_cpu_baseline = True
    K = max(args.steps, 1)
    wl = args.workload
    kind = WORKLOADS[wl]["kind"]
    N = args.size or WORKLOADS[wl]["size"]
    metric = "IMPES Gcell-updates/s" if kind == "impes" else "MUSCL transport Gcell-updates/s"
    config = common_config(wl, N, args.eps)

    # ------------------------------------------------------------------ reference arm (CPU; rank 0 alone)
    if args.impl == "reference":
        if rank != 0:
            return 0
        val, ms, base, _ = cpu_arm(wl, N, K, W, args.eps)
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "Gcell-updates/s", "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import numpy as np
    import torch
    import torch.distributed as dist
    import impec2d_b200 as I

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    if kind == "transport" and world > 1:
        raise SystemExit("bench.py: hyperbolic_test (periodic test mode) runs on one GPU, as BASELINE config 2 names it")
    need_gb = 27.0 * 8.0 * N * N / world / 1e9
    if need_gb > 165.0:
        raise SystemExit("bench.py: %s at %d^2 needs ~%.0f GB per GPU on %d GPU(s); use more GPUs or --size" % (wl, N, need_gb, world))
    torch.cuda.set_device(local_rank)
    host_binding = None
    if world > 1 and os.environ.get("IMPEC_BENCH_BIND", "1") != "0":
        # one process per GPU: keep this rank's host threads, its pinned buffers (allocated below) and the library's staging buffers on the
        # NUMA node its GPU hangs off -- with 8 ranks half of the host <-> device traffic of the end-to-end arm would otherwise cross the
        # socket interconnect.  Not at N = 1: the cpu_baseline leg of that line uses every host thread of the box.
        try:
            ncpu, cpus = I.bind_host_to_device(local_rank)
            host_binding = {"cpus": ncpu, "cpulist": cpus} if ncpu else "unavailable (topology not visible to this process): unbound"
        except Exception as e:                               # never let placement stop a measurement
            host_binding = "failed: %s" % e
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                  # NCCL prints its version banner on stdout: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # a dedicated (non-default) stream shared by the library, the timing events and torch's collectives: the library keeps a NULL
    # stream argument for "create my own", so the legacy default stream cannot be handed to it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    P = workload_params(wl, N, args.eps, I)
    # one process per GPU; rank r owns rows [r N/P, (r+1) N/P) (slab decomposition along x).  The NCCL communicator of
    # the library is bootstrapped with an id broadcast over torch.distributed.
    nccl_id = None
    if world > 1:
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            rc = I.load().impec_nccl_unique_id(idbuf)
            assert rc == 0, "impec_nccl_unique_id failed (%d)" % rc
        idt = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, src=0)
        nccl_id = bytes(idt.cpu().tolist())
    ctx = I.Context(P, device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id, stream=C.c_void_p(stream.cuda_stream))
    options = {}
    for kv in args.opt:
        k, v = kv.split("="); ctx.set_option(k, float(v)); options[k] = float(v)
    row0, nrows = ctx.local_rows()
    # host inputs in pinned memory (reference layouts, this rank's rows)
    shape = (nrows + 2 * P.g, N + 2 * P.g)
    c_host_t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    q_host_t = torch.empty(nrows * N, dtype=torch.float64, pin_memory=True)
    c_host = c_host_t.numpy(); q_host = q_host_t.numpy()
    c_host[...] = 0.0
    if kind == "transport":
        c_host[P.g:-P.g, P.g:-P.g] = sine_cell_averages(N)
        q_host[...] = 0.0
        vel = np.empty(ctx.velocity_shape()); vel[:, :, 0] = -1.0; vel[:, :, 1] = 1.0      # CalcAdvectionVelocities, src/Impec2DSolver.cpp:1003-1015
        ctx.upload_velocity(vel); del vel
    else:
        I.synth_state_rows(P, row0, nrows, 20260917, out=c_host)
        q_host[...] = I.init_source_rows(P, row0, nrows, 1.0, 0.05, 0)
    ctx.upload_c_rows(c_host); ctx.upload_q_rows(q_host)
    ctx.synchronize()
    dt_fixed = ctx.max_stable_dt() if kind == "transport" else None           # dt_max is computed once, before the loop (src/Impec2DSolver.cpp:908)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"t": 0.0}
    its, relres, true_rr = [], [], []

    def one_step(record):
        if kind == "transport":
            ctx.transport_step(dt_fixed)
            return
        dt, info = ctx.impes_step(state["t"], 0.0); state["t"] += dt
        if record:
            its.append(info.iterations); relres.append(info.relres); true_rr.append(info.true_relres)

    for _ in range(W):
        one_step(False)

    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.25)
    barrier()
    l0 = ctx.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        one_step(True)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - l0
    clocks = sampler.finish()
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms_total = float(tt.item())
    cells_all = N * N * K                           # whole job: the grid is cut into slabs, total work is fixed
    value = cells_all / (ms_total * 1e-3) / 1e9

    # ------------------------------------------------------------------ per-kernel device times (CUDA events, live)
    prof = {}
    if kind == "impes":
        names = (("vcycle", 5), ("fused_pre", 10), ("fused_post", 10), ("fused_plain", 10), ("pcg_matvec", 10), ("update_xr", 10), ("update_r", 10), ("update_xxr", 10),
                 ("sweep_x", 10), ("sweep_y", 10), ("build_operator", 5), ("calc_velocities", 5), ("transport_step", 3))
    else:
        names = (("sweep_x", 10), ("sweep_y", 10), ("transport_step", 5))
    if not args.profile_mode:
        if kind == "impes" and dt_fixed is None:
            ctx.set_option("profile_dt", ctx.max_stable_dt())
        elif dt_fixed is not None:
            ctx.set_option("profile_dt", dt_fixed)
        for name, reps in names:
            prof[name] = ctx.profile_kernel(name, reps)
    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "workload": wl, "ms_per_step": ms_total / K, "pcg_iterations": its, "gpu_launches": int(launches)}))
        ctx.close()
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # Algorithmic bytes per fine-level launch (FP64, each distinct array once):
    #   k_mg_march_post<PROLONG> : R x,f (16) + FP32 AW,AS,AV,DINV (16) + coarse e (2) + W x (8)  = 42 B/cell
    #   k_mg_march_pre           : R f (8) + FP32 AW,AS,AV,DINV (16) + W x (8) + coarse f (2)     = 34 B/cell
    #   k_mg_march_post<PLAIN>   : R x,f (16) + FP32 coefficients (16) + W x (8)                   = 40 B/cell
    #   k_pcg_matvec_march       : R z,d,AW,AS,AV (40) + W d,s (16)                                = 56 B/cell
    #   k_update_xr 48;  k_muscl_x / k_muscl_y: R c,vel + W c = 24 B/cell
    cells = float(N) * N / world                   # cells per GPU: the kernels are timed on this rank's slab
    #   k_update_xr<0> 48; the solve runs <1> (r only, 24) and <2> (x with both pending directions + r, 56) alternately: 80 per two iterations
    kb = {"fused_post": 42.0, "fused_pre": 34.0, "fused_plain": 40.0, "pcg_matvec": 56.0, "update_xr": 48.0, "update_r": 24.0, "update_xxr": 56.0,
          "sweep_x": 24.0, "sweep_y": 24.0}
    per_kernel = {k: {"ms": prof[k], "algorithmic_bytes": kb[k] * cells, "achieved_gbs": kb[k] * cells / (prof[k] * 1e-3) / 1e9,
                      "frac": kb[k] * cells / (prof[k] * 1e-3) / 1e9 / peak} for k in kb if k in prof}
    step_s = ms_total / K * 1e-3 * world           # GPU-seconds per step
    traffic = None; traffic_src = None             # measured DRAM bytes per launch of the dominant kernel (ncu capture, 1 GPU, same grid)
    if kind == "impes":
        dom = "fused_post"; dom_name = "k_mg_march_post<PROLONG> (prolong + correct + red-black post-smooth + <r,z>, fine level; register march, L2 prefetch 4 rows ahead)"
    else:
        # the slower of the two sweep kernels (each runs twice per step)
        dom = "sweep_x" if prof["sweep_x"] >= prof["sweep_y"] else "sweep_y"
        dom_name = ("k_muscl_x (MUSCL-Hancock x sweep: a thread marches down a column, fluxes on the fly; 2 of the 4 sweeps of a step)" if dom == "sweep_x" else
                    "k_muscl_y2 (MUSCL-Hancock y sweep: column pairs per lane, neighbours by shuffle, fluxes on the fly; 2 of the 4 sweeps of a step)")
    try:
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
        for f in reversed(cands):
            tr = json.load(open(os.path.join(ROOT, "profiles", f)))
            if tr.get("grid") == [N, N] and world == 1 and tr.get(dom):
                traffic = tr.get(dom); traffic_src = "profiles/%s (ncu --set full capture at commit %s; NOT measured in this run)" % (f, tr.get("commit", "?"))
                break
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_kernel[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_kernel[dom]["algorithmic_bytes"], "kernels": per_kernel, "kernel_ms": prof}
    if kind == "impes":
        # What the fused cycle actually moves (per FINE cell): level 0 and 1 PRE 34 + POST 42; W-cycle levels 2-4
        # PRE 34 + POST 42 + MID 42 + POST 42 = 160 per visit, visited 1 / 2 / 4 times; below that 76 per visit, 8 visits:
        #   76 (1 + 1/4) + 160 (1/16 + 2/64 + 4/256) + 76 * 8 (1/1024 + 1/4096 + ...) = 113.3 B per fine cell
        # SURVEY 8d's model of an UNFUSED cycle with FP64 coefficients is 208 B/cell: the fused FP32-coefficient cycle moves fewer bytes,
        # so fractions against that model can exceed 1 and are reported only for reference; frac_moved is the efficiency measure.
        w_from, w_to = ctx.mg_w_range()
        cyc_b = cycle_bytes_per_fine_cell(w_from, w_to, ctx.mg_levels())         # 113.3 for the default range 2..4 on 13 levels
        vc_moved = cyc_b * cells; vc_model = 208.0 * cells
        its_mean = float(sum(its)) / max(len(its), 1)
        NN = float(N) * N
        cn = 0.0
        if WORKLOADS[wl]["diffusion"]:
            # per CN half step: links 32 + DINV/FP32 copies 56 + rhs 40 + n_cn (k_sgs_march 34 + mat-vec 56 + update 48) + copy-out 16; two per step
            cn = 2.0 * (32.0 + 56.0 + 40.0 + 9.0 * (34.0 + 56.0 + 40.0) + 16.0)
        step_model = (32.0 + its_mean * (208.0 + 128.0) + 48.0 + 128.0) * NN
        # operator build 74 (one march c -> links, DINV, FP32 copies: 48; coarsening of the links down the hierarchy: 26), per PCG iteration the
        # cycle 113.3 + mat-vec 56 + update 40 (24 and 56 alternately: deferred solution update), velocity + CFL 32 (+ 8 ghost fills), transport 4 x 24 (+ 8 source / fills)
        step_moved = (ctx_option_bytes_operator() + its_mean * (cyc_b + 56.0 + 40.0) + 40.0 + 104.0 + cn) * NN
        roofline["vcycle"] = {"ms": prof["vcycle"], "moved_bytes": vc_moved, "achieved_moved": vc_moved / (prof["vcycle"] * 1e-3) / 1e9,
                              "frac_moved": vc_moved / (prof["vcycle"] * 1e-3) / 1e9 / peak,
                              "survey_model_bytes": vc_model, "frac_survey_model": vc_model / (prof["vcycle"] * 1e-3) / 1e9 / peak}
        roofline["step"] = {"moved_bytes": step_moved, "achieved_moved": step_moved / step_s / 1e9, "frac_moved": step_moved / step_s / 1e9 / peak,
                            "survey_model_bytes": step_model, "frac_survey_model": step_model / step_s / 1e9 / peak,
                            "note": "frac_moved = bytes the implementation reads+writes / time / peak (the efficiency measure); the survey model assumes an unfused FP64-coefficient cycle"}
    else:
        step_moved = 4.0 * 24.0 * float(N) * N
        roofline["step"] = {"moved_bytes": step_moved, "achieved_moved": step_moved / step_s / 1e9, "frac_moved": step_moved / step_s / 1e9 / peak,
                            "survey_model_bytes": 128.0 * float(N) * N, "frac_survey_model": 128.0 * float(N) * N / step_s / 1e9 / peak}

    # ------------------------------------------------------------------ end-to-end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        ctx.download_c_rows(c_host)
        h2d0 = ctx.bytes_h2d(); d2h0 = ctx.bytes_d2h()
        barrier()
        t0 = time.perf_counter()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(K):
            ctx.upload_c_rows(c_host); ctx.upload_q_rows(q_host)
            one_step(False)
            ctx.download_c_rows(c_host)
        f1.record(stream)
        barrier()
        ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
        h2d = (ctx.bytes_h2d() - h2d0) // K; d2h = (ctx.bytes_d2h() - d2h0) // K
        if world > 1:
            tt = torch.tensor([ms_e2e, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt[:1], op=dist.ReduceOp.MAX); dist.all_reduce(tt[1:], op=dist.ReduceOp.SUM)
            ms_e2e = float(tt[0].item()); h2d = int(tt[1].item()); d2h = int(tt[2].item())
        e2e = {"value": cells_all / (ms_e2e * 1e-3) / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e2e / K,
               "host_input_bytes_per_step": int(c_host.nbytes + q_host.nbytes) * world,
               "note": "host inputs per step: c (ghosted) and the dense q array; q is scanned on the host while c is in flight and only its non-zero "
                       "(well) cells cross PCIe, so h2d_bytes_per_step (counted inside the library from the copies it issues) is below host_input_bytes_per_step"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, cpu_baseline, _ = cpu_arm(wl, N, 2, 1, args.eps, CPU_BUDGET_BASELINE_S)

    if rank == 0:
        run = {"mg_levels": ctx.mg_levels(), "mg_w_range": list(ctx.mg_w_range()), "comm_backend": ctx.comm_backend(),
               "parallelism": "1 GPU" if world == 1 else "%d slabs of %d rows along x; halo rows and scalar reductions over %s; multigrid levels below 4M cells replicated after an all-gather" % (world, nrows, ctx.comm_backend()),
               "coefficients": "multigrid cycle reads FP32 copies of the link/diagonal arrays (all arithmetic FP64; PCG mat-vec uses the FP64 operator)"}
        run["halo_exchanges_overlapped_with_interior_chunks"] = int(ctx.overlapped_exchanges())
        if world > 1:
            run["host_binding_rank0"] = host_binding           # impec_bind_host_to_device: CPUs local to the rank's GPU
        if options:
            run["options"] = options                       # non-default library options of this run (A/B measurements)
        if kind == "impes":
            run.update({"pcg_iterations": its, "relres": relres, "true_relres": true_rr})
        line = {"metric": metric, "value": value, "unit": "Gcell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "run": run,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline}
        if kind == "impes":
            line["vcycle_ms"] = prof["vcycle"]
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def cycle_bytes_per_fine_cell(w_from, w_to, nlevels):
    """Bytes one multigrid cycle reads + writes per FINE cell: a level visited once per visit of its parent moves
    PRE 34 + POST 42 = 76 B per cell of its own, a level of the W range (visits the next level twice) PRE 34 + POST 42 + MID 42 + POST 42 = 160;
    level l has 4^-l of the fine cells."""
    total, visits = 0.0, 1.0
    for l in range(nlevels):
        w = w_from <= l <= w_to
        total += visits * (160.0 if w else 76.0) / 4.0 ** l
        if w:
            visits *= 2.0
    return total


def ctx_option_bytes_operator():
    """Bytes per cell the operator build moves: the fused build reads c and writes the FP64 links, DINV and the FP32
    copies (48), then the links are coarsened level by level (26, measured: 1.1 ms at 16384^2)."""
    return 74.0


if __name__ == "__main__":
    sys.exit(main())
