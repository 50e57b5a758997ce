# where the end-to-end (host buffer) overhead of a step goes: upload c, upload q (host scan + compact copy), download c, at the bench size
import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import impec2d_b200 as I
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
ctx = I.Context(P)
c_t = torch.empty((N + 4, N + 4), dtype=torch.float64, pin_memory=True); q_t = torch.empty(N * N, dtype=torch.float64, pin_memory=True)
c = c_t.numpy(); q = q_t.numpy()
I.synth_state_rows(P, 0, N, 20260917, out=c); q[...] = I.init_source_rows(P, 0, N, 1.0, 0.05, 0)
def timed(name, fn, reps=5):
    fn(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    ctx.synchronize()
    print("%-28s %8.2f ms" % (name, (time.perf_counter() - t0) / reps * 1e3), flush=True)
timed("upload_c_rows (2.15 GB)", lambda: ctx.upload_c_rows(c))
timed("upload_q_rows (scan 2.15 GB)", lambda: ctx.upload_q_rows(q))
timed("upload c + q", lambda: (ctx.upload_c_rows(c), ctx.upload_q_rows(q)))
timed("download_c_rows (2.15 GB)", lambda: ctx.download_c_rows(c))
ctx.close()
