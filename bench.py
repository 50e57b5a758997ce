#!/usr/bin/env python
"""bench.py -- IMPES hot-path benchmark for the B200-native Impec2D path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--eps E] [--impl ours|reference]

Workload (BASELINE.json metric: "IMPES Gcell-updates/s & MG V-cycle ms at 16384^2"):
  parallel.cfg operator (9-point blend nu=2/3, harmonic mobility, mu_res/mu_inj = 41, quarter-circle wells,
  MUSCL symmetric-parallel transport, CFL 0.5) on an S x S grid (default 16384) started from the developed
  synthetic fingering state of SURVEY.md 8(d).  One step = one iteration of Impec2DSolver::Solve's loop:
  FillBoundaries, MG-PCG pressure solve from x0 = 0 to the cfg tolerance (1e-14, HYPRE's recurrence-residual
  test), CalcVelocities, MaxStableTimestep, transport step.  All FP64.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same metric
through the C ABI with pinned HOST buffers (c and q uploaded, c downloaded, every step).
`--impl reference` times the CPU restatement of the reference's path (oracle/, all host threads) on a bounded
sample of the same workload: the reference itself cannot be built in this image (MPI, HYPRE, blitz++,
libconfig++ are absent).
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
CPU_SAMPLE_N = 4096          # grid of the bounded CPU sample (same operator, same state generator): ~4.5 s per step on 16 cores


def workload_params(N, eps, module):
    return module.default_params(N1=N, N2=N, x_ub=SQRT2, y_ub=SQRT2, stype=0, nu=2.0 / 3.0, mu_res=41.0, mu_inj=1.0,
                                 avgproc=1, cfl=0.5, eps=eps, hyper_method=0, dim_splitting=4, diffusion=0)


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


def cpu_reference_run(steps, warmup, eps, n=CPU_SAMPLE_N):
    """The oracle (CPU restatement of the reference path) on all host threads: `steps` IMPES steps on an n x n
    sample of the workload.  Returns (Gcell-updates/s, ms per step, cores, iterations)."""
    import numpy as np
    import oracle as O
    lib = O.lib()
    P = workload_params(n, eps, O)
    cores = lib.orc_num_threads()
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
    return n * n * steps / el / 1e9, el / steps * 1e3, cores, its


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--eps", type=float, default=1e-14, help="pressure tolerance (examples/parallel.cfg:52)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="for runs under ncu: honour --warmup as given, skip e2e / cpu baseline / kernel timing (numbers printed in this mode are not bench values)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = args.warmup if args.profile_mode else max(args.warmup, 3)
    if args.profile_mode:
        args.no_e2e = True; args.no_cpu_baseline = True
    K = max(args.steps, 1)
    N = args.size
    metric = "IMPES Gcell-updates/s"
    workload = "parallel.cfg IMPES step (MG-PCG pressure to eps + velocity + CFL + MUSCL transport) at %d^2, developed synthetic state" % N

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        val, ms, cores, its = cpu_reference_run(K, min(W, 1), args.eps)
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "Gcell-updates/s", "n_gpus": args.gpus, "steps": K,
                "warmup": min(W, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "eps": args.eps, "pcg_iterations": its},
                "cpu_baseline": {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port",
                                 "sample": "%d IMPES steps on a %d^2 sample of the workload (same operator, eps, state generator); "
                                           "CPU restatement of the reference path, OpenMP over all host threads" % (K, CPU_SAMPLE_N)},
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
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()

    P = workload_params(N, args.eps, I)
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
    row0, nrows = ctx.local_rows()
    # host inputs in pinned memory (reference layouts, this rank's rows)
    shape = (nrows + 2 * P.g, N + 2 * P.g)
    c_host_t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    q_host_t = torch.empty(nrows * N, dtype=torch.float64, pin_memory=True)
    c_host = c_host_t.numpy(); q_host = q_host_t.numpy()
    c_host[...] = 0.0
    I.synth_state_rows(P, row0, nrows, 20260917, out=c_host)
    q_host[...] = I.init_source_rows(P, row0, nrows, 1.0, 0.05, 0)
    ctx.upload_c_rows(c_host); ctx.upload_q_rows(q_host)
    ctx.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t = 0.0
    for _ in range(W):
        dt, info = ctx.impes_step(t, 0.0); t += dt

    sampler = ClockSampler(local_rank); sampler.start()
    time.sleep(0.25)
    its, relres, true_rr = [], [], []
    barrier()
    l0 = ctx.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        dt, info = ctx.impes_step(t, 0.0); t += dt
        its.append(info.iterations); relres.append(info.relres); true_rr.append(info.true_relres)
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
    for name, reps in () if args.profile_mode else (("vcycle", 5), ("fused_pre", 10), ("fused_post", 10), ("fused_plain", 10), ("smooth_sweep", 10), ("residual", 10), ("restrict", 10), ("prolong_add", 10), ("pcg_matvec", 10), ("matvec_dot", 10),
                       ("update_xr", 10), ("dot", 10), ("update_d", 10), ("sweep_x", 10), ("sweep_y", 10), ("build_operator", 5), ("calc_velocities", 5)):
        prof[name] = ctx.profile_kernel(name, reps)
    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step": ms_total / K, "pcg_iterations": its, "gpu_launches": int(launches)}))
        ctx.close()
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # Algorithmic bytes per fine-level launch (FP64, each distinct array once; DESIGN.md "kernels"):
    #   k_mg_fused<POST>  : R x,f (16) + FP32 AW,AS,AV,DINV (16) + coarse e (2) + W x (8)   = 42 B/cell
    #   k_mg_fused<PRE>   : R f (8) + FP32 AW,AS,AV,DINV (16) + W x (8) + coarse f (2)     = 34 B/cell
    #   k_mg_fused<PLAIN> : R x,f (16) + FP32 coefficients (16) + W x (8)                   = 40 B/cell
    #   k_pcg_matvec_march: R z,d,AW,AS,AV (40) + W d,s (16)                        = 56 B/cell
    #   k_update_xr 48, k_muscl_x / k_muscl_y: R c,vel + W c = 24 B/cell
    #   one multigrid cycle, all levels: SURVEY 8d's model figure (unfused, FP64 coefficients) = 208 B/cell
    cells = float(N) * N / world                   # cells per GPU: the kernels are timed on this rank's slab
    kb = {"fused_post": 42.0, "fused_pre": 34.0, "fused_plain": 40.0, "pcg_matvec": 56.0, "matvec_dot": 40.0, "update_xr": 48.0, "update_d": 24.0,
          "sweep_x": 24.0, "sweep_y": 24.0, "residual": 48.0, "smooth_sweep": 48.0}
    per_kernel = {k: {"ms": prof[k], "algorithmic_bytes": kb[k] * cells, "achieved_gbs": kb[k] * cells / (prof[k] * 1e-3) / 1e9,
                      "frac": kb[k] * cells / (prof[k] * 1e-3) / 1e9 / peak} for k in kb}
    dom = "fused_post"
    traffic = None                                 # measured DRAM bytes of the dominant kernel from the committed ncu capture (same grid, 1 GPU)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1q_traffic.json")))
        if tr.get("grid") == [N, N] and world == 1:
            traffic = tr.get(dom)
    except Exception:
        pass
    vc_bytes = 208.0 * float(N) * N / world
    # What the fused cycle actually moves (per FINE cell): level 0 and 1 PRE 34 + POST 42; W-cycle levels 2-4
    # PRE 34 + POST 42 + MID 42 + POST 42 = 160 per visit, visited 1 / 2 / 4 times; below that 76 per visit, 8 visits:
    #   76 (1 + 1/4) + 160 (1/16 + 2/64 + 4/256) + 76 * 8 (1/1024 + 1/4096 + ...) = 113.3 B per fine cell
    vc_moved = 113.3 * float(N) * N / world
    its_mean = float(sum(its)) / max(len(its), 1)
    # one IMPES step: SURVEY 8d model = 32 + n_it (208 + 128) + 48 + 128 B/cell; bytes the implementation moves =
    # operator build 168 + n_it (cycle 113.3 + mat-vec 56 + update 48) + velocity/CFL 56 + transport 4 x 24 + fills
    step_model = (32.0 + its_mean * (208.0 + 128.0) + 48.0 + 128.0) * float(N) * N
    step_moved = (168.0 + its_mean * (113.3 + 56.0 + 48.0) + 56.0 + 96.0) * float(N) * N
    step_s = ms_total / K * 1e-3 * world           # GPU-seconds per step
    roofline = {"bound": "hbm", "kernel": "k_mg_march_post<PROLONG> (prolong + correct + red-black post-smooth + <r,z>, fine level; register march)",
                "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": per_kernel[dom]["frac"], "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": per_kernel[dom]["algorithmic_bytes"],
                "vcycle": {"ms": prof["vcycle"], "algorithmic_bytes": vc_bytes, "achieved": vc_bytes / (prof["vcycle"] * 1e-3) / 1e9,
                           "frac": vc_bytes / (prof["vcycle"] * 1e-3) / 1e9 / peak,
                           "note": "algorithmic_bytes = SURVEY 8d model (unfused cycle, FP64 coefficients: 208 B/cell); moved_bytes = what the fused cycle reads and writes (113.3 B/cell)",
                           "moved_bytes": vc_moved, "frac_moved": vc_moved / (prof["vcycle"] * 1e-3) / 1e9 / peak},
                "step": {"survey_model_bytes": step_model, "frac_survey_model": step_model / step_s / 1e9 / peak,
                         "moved_bytes": step_moved, "achieved_moved": step_moved / step_s / 1e9, "frac_moved": step_moved / step_s / 1e9 / peak},
                "kernels": per_kernel, "kernel_ms": prof}

    # ------------------------------------------------------------------ end-to-end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        ctx.download_c_rows(c_host)
        barrier()
        t0 = time.perf_counter()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(K):
            ctx.upload_c_rows(c_host); ctx.upload_q_rows(q_host)
            dt, info = ctx.impes_step(t, 0.0); t += dt
            ctx.download_c_rows(c_host)
        f1.record(stream)
        barrier()
        ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
        if world > 1:
            tt = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms_e2e = float(tt.item())
        e2e = {"value": cells_all / (ms_e2e * 1e-3) / 1e9, "unit": "Gcell-updates/s",
               "h2d_bytes_per_step": int(c_host.nbytes + q_host.nbytes) * world, "d2h_bytes_per_step": int(c_host.nbytes) * world, "ms_per_step": ms_e2e / K}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, ms, cores, cits = cpu_reference_run(2, 1, args.eps)
        cpu_baseline = {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port", "ms_per_step": ms, "pcg_iterations": cits,
                        "sample": "2 IMPES steps on a %d^2 sample of the workload (same operator, eps, state generator) after 1 warm-up; "
                                  "CPU restatement of the reference path (oracle/), OpenMP over all host threads" % CPU_SAMPLE_N}

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "Gcell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "grid": [N, N], "eps": args.eps, "stop_rule": "<r,r>/<b,b> < eps^2, recurrence residual, x0 = 0 (HYPRE StructPCG semantics)",
                           "pcg_iterations": its, "relres": relres, "true_relres": true_rr, "mg_levels": ctx.mg_levels(),
                           "parallelism": "1 GPU" if world == 1 else "%d slabs of %d rows (halo send/recv over NCCL; multigrid levels below %d rows per slab replicated after an all-gather)" % (world, nrows, 8),
                           "coefficients": "multigrid cycle reads FP32 copies of the link/diagonal arrays (all arithmetic FP64; PCG mat-vec uses the FP64 operator)",
                           "l2": "inputs larger than L2 (every field is %.2f GB)" % (8.0 * N * N / 1e9)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "vcycle_ms": prof["vcycle"]}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
