# side measurement under torchrun: cost of the communication primitives and of the replicated multigrid tail
import os, sys, time, json, ctypes as C
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import impec2d_b200 as I
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.cuda.set_device(lr)
nccl_id = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    idbuf = C.create_string_buffer(128)
    if rank == 0: assert I.load().impec_nccl_unique_id(idbuf) == 0
    idt = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device="cuda"); dist.broadcast(idt, src=0)
    nccl_id = bytes(idt.cpu().tolist())
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
ctx = I.Context(P, device=lr, rank=rank, nranks=world, nccl_id=nccl_id, stream=C.c_void_p(stream.cuda_stream))
row0, nrows = ctx.local_rows()
import numpy as np
c = np.zeros((nrows + 2 * P.g, N + 2 * P.g)); I.synth_state_rows(P, row0, nrows, 20260917, out=c)
q = I.init_source_rows(P, row0, nrows, 1.0, 0.05, 0)
ctx.upload_c_rows(c); ctx.upload_q_rows(q)
t = 0.0
for _ in range(2): dt, info = ctx.impes_step(t, 0.0); t += dt
out = {"world": world, "its": info.iterations, "backend": ctx.comm_backend()}
if rank == 0: print("backend", ctx.comm_backend(), flush=True)
for name, reps in (("halo_x2", 200), ("halo_f3", 200), ("allreduce1", 200), ("gather_rhs", 100), ("tail", 50), ("cycle@1", 20), ("cycle@2", 20), ("cycle@3", 20), ("cycle@4", 20), ("cycle@5", 20), ("fused_pre@1", 20), ("fused_post@1", 20), ("fused_pre@3", 20), ("fused_post@3", 20), ("vcycle", 10),
                   ("fused_pre", 10), ("fused_post", 10), ("pcg_matvec", 10), ("update_xr", 10)):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms = ctx.profile_kernel(name, reps)
    wall = (time.perf_counter() - t0) / (reps + 1) * 1e3
    out[name] = {"device_ms": ms, "wall_ms": wall}
    if rank == 0: print("%-12s device %.4f ms  wall %.4f ms" % (name, ms, wall), flush=True)
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): dt, info = ctx.impes_step(t, 0.0); t += dt
torch.cuda.synchronize(); out["ms_per_step"] = (time.perf_counter() - t0) / 3 * 1e3
if rank == 0:
    print("step %.1f ms its %d" % (out["ms_per_step"], info.iterations))
    json.dump(out, open('gpurun_out/comm_diag_n%d.json' % world, 'w'), indent=1)
ctx.close()
if world > 1: dist.destroy_process_group()
