// impec2d_b200.cu -- context, host orchestration and the C ABI (include/impec2d_b200.h).
// Unity build: the kernels are included below so that nvcc (product) and g++ -DIMPEC_EMU (CPU
// emulation for tests) compile exactly the same source.
#include "common.h"
#include "kernels_transport.inl"
#include "kernels_pressure.inl"
#include "kernels_build.inl"
#include "kernels_fused.inl"
#include "kernels_march.inl"
#include "kernels_small.inl"
#include "../../include/impec2d_b200.h"

#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
#include <nccl.h>
#endif
#include <string>
#include <vector>
#include <cstdio>
#include <cstdarg>
#include <climits>
#include <sched.h>
#include <algorithm>
#include <thread>
#include <atomic>
#include <functional>

#define BX 32
#define BY 8
#include "comm_p2p.inl"
#ifdef IMPEC_EMU
#include "emu_filecomm.inl"
#endif

// ------------------------------------------------------------------------------------------------
struct Level {
    Grid g;
    double *AW = nullptr, *AS = nullptr, *AV = nullptr, *DINV = nullptr;
    float *AWf = nullptr, *ASf = nullptr, *AVf = nullptr, *DINVf = nullptr;      // FP32 copies for the cycle kernels
    double *x = nullptr, *f = nullptr, *t = nullptr, *r = nullptr;
    long long pin = -1, pin_end = -1;       // Dirichlet cells of the fine level: index range [pin, pin_end) (one cell, or one row for source_type horizontal)
    long long sig_idx = -1, sig_end = -1;   // coarse levels: cells carrying the pins' diagonal shift
    long long sig_off = 0;                  // offset of this level's shifts in impec_ctx::sigma_dev
    bool dist = false;           // rows of this level are slab-distributed over the ranks
    int th = 64;                 // tile height of the fused kernels on this level (64 / 32 / 16)
    int march_rows = 0;          // > 0: the level runs the marching kernels (kernels_march.inl) with chunks of this many rows
    // first replicated level only: this rank's slab view of it and slab-sized staging for the all-gathers
    Grid gl; double *AWl = nullptr, *ASl = nullptr, *AVl = nullptr, *fl = nullptr;
};

struct impec_ctx {
    impec_params P;
    DevParams D;
    Grid g;
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // fields (padded layout)
    double *c = nullptr, *p = nullptr, *q = nullptr, *u = nullptr, *v = nullptr;
    double *ta = nullptr, *tb = nullptr;             // transport temporaries (contiguous pair)
    double *MU = nullptr, *KV11 = nullptr, *KV22 = nullptr;
    double *r = nullptr, *z = nullptr, *d = nullptr, *s = nullptr, *rhs = nullptr, *t0 = nullptr, *d2buf = nullptr;
    std::vector<Level> L;
    double *sigma_dev = nullptr;                    // diagonal shifts of the pins on the coarse levels (Level::sig_off)
    long long sigma_len = 0;
    // reductions
    double *part = nullptr; long long part_cap = 0;
    double *scal = nullptr;                         // device scalars
    double *h_scal = nullptr;                       // pinned host mirror
    // wells
    int nwell = 0, well_cap = 0; int *well_i = nullptr, *well_j = nullptr; double *well_q = nullptr, *well_md = nullptr;
    int *hw_i = nullptr, *hw_j = nullptr; double *hw_q = nullptr; long long hw_cap = 0;    // pinned staging of the compact well list (scan_wells)
    cudaEvent_t ev_wells = nullptr;                 // recorded after the copies out of the staging buffers
    double q_pin = 0.0; bool have_q = false;
    int prev_first_sweep = 1;                       // Y_SWEEP (src/HyperbolicSolver.cpp:84)
    bool mob_valid = false, op_valid = false;
    bool warm_start = false;
    bool fused = true;                              // fused tile kernels (kernels_fused.inl) vs one kernel per colour pass
    bool stagnation_guard = false;                  // stop PCG once the TRUE residual has hit its floating-point floor (off: reference semantics)
    bool links_f32 = true;                          // cycle kernels read FP32 copies of the coefficients (FP64 arithmetic)
    bool fast_r0 = true;                            // initial PCG residual for x0 = 0: copy of rhs + fix-up around the pin instead of a residual pass
    bool sgs_march = true;                          // Crank-Nicolson preconditioner (one symmetric sweep from zero) as ONE march (k_sgs_march)
    bool march = true;                              // big levels: PRE / POST / PLAIN as register marches (kernels_march.inl); 0 = shared-memory tiles
    bool fused_mid = true;                          // W-cycle revisit: sweep + residual + restriction in one tile kernel (FUSED_MID)
    bool matvec_march = true;                       // PCG direction update + mat-vec as a register march (k_pcg_matvec_march); 0 = shared-memory tiles
    bool sweep_shared = true;                       // MUSCL sweeps evaluate each reconstruction once (k_muscl_x / k_muscl_y); 0 = k_sweep
    bool vel_march = true;                          // production runs: velocity + CFL reductions from c and p in ONE march (k_velocity_march); 0 = k_velocity + k_dt_reduce
    bool dt_cached = false;                         // scal[S_DT], scal[S_DT+1] hold the reductions of the current velocity field
    bool build_fused = true;                        // production runs: c -> links, DINV and FP32 copies in ONE march (kernels_build.inl); 0 = the five-pass build
    bool use_small = true;                          // levels of <= SM_MAXCELLS cells: whole V-cycle in one single-CTA launch (kernels_small.inl)
    long long launches = 0;
    impec_solve_info last_para = {0, 0., 0., 0};     // the most recent Crank-Nicolson solve inside impec_transport_step (impec_last_parabolic_info)
    unsigned long long bytes_h2d = 0, bytes_d2h = 0;   // bytes of the field transfers this context issued (impec_transfer_bytes)
    double profile_dt = 0.0;                        // time step used by the "transport_step" / "diffusion_split" profiling hooks
    long long march_launches = 0;                   // launches of the marching multigrid kernels (diagnostic, IMPEC_VERBOSE)
    std::string err;
    std::vector<void *> allocs;
    // multi-GPU: slab decomposition along i
    int Ld = 0;                                     // levels [0,Ld) are distributed, [Ld,nl) replicated on every rank
    int N1g = 0;                                    // global number of rows
    impec_comm_callbacks cb; bool has_cb = false;
#ifdef IMPEC_EMU
    EmuFileComm *filecomm = nullptr;                // emulated multi-process runs without a Python harness (emu_filecomm.inl)
#endif
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    ncclComm_t nccl = nullptr;
#endif
    bool comm_ready = false;
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    // peer-memory transport (comm_p2p.inl): my mailbox, the mapped mailboxes of the other ranks, exchange epochs
    bool p2p = false;                               // mailboxes mapped: scalar all-reduces go through peer memory
    bool p2p_halo = false;                          // ... and so do the halo rows (IMPEC_P2P_HALO=0/1)
    unsigned char *mail = nullptr, *peer_mail[P2P_MAXR] = {};
    size_t p2p_slot_bytes = 0;
    size_t p2p_win_bytes = 0;                       // one all-gather window (the restricted right-hand side of the first replicated level)
    bool p2p_gather = false;                        // ... and that all-gather too (IMPEC_P2P_GATHER=0/1)
    unsigned long long halo_epoch = 0, ar_epoch = 0, ag_epoch = 0;
#endif
    // Halo exchange overlapped with interior compute: the exchange and the two slab-edge row chunks of its consumer run on a
    // second, high-priority stream while the interior chunks of the same kernel run on the main stream (halo_overlapped()).
    // marching kernels: rows of L2 software prefetch ahead of the march, per kernel family (PRE, POST / PLAIN, mat-vec, CN sweep);
    // option "l2_prefetch" sets all four, IMPEC_L2_PREFETCH="pre,post,mv,sgs" (or one number) each.  0 = off.
    int pf_pre = 4, pf_post = 4, pf_mv = 4, pf_sgs = 4;
    int pf_sweep = 4;                               // MUSCL sweeps (option "sweep_prefetch" / IMPEC_SWEEP_PREFETCH)
    bool defer_x = true;                            // PCG: merge the solution updates of two iterations (k_update_xr MODE 1 / 2; option "defer_x")
    bool sweep_y_pairs = true;                      // MUSCL y sweep: column pairs per lane + shuffles (k_muscl_y2); 0 = the shared-memory tile kernel k_muscl_y
    bool overlap = true;                            // option "halo_overlap" / IMPEC_HALO_OVERLAP=0/1
    long long overlapped = 0;                       // exchanges issued that way (diagnostic)
#ifndef IMPEC_EMU
    cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
#endif
    // CUDA graph of the replicated multigrid tail (levels >= tail_level): identical launch sequence every visit
    int tail_level = -1;
    bool use_graph = true, graph_ready = false;
    long long graph_launches = 0;
    std::vector<std::pair<double *, double *> > tail_init, tail_final;
#ifndef IMPEC_EMU
    cudaGraph_t tail_graph = nullptr; cudaGraphExec_t tail_exec = nullptr;
#endif
};

enum { S_GAMMA0 = 0, S_SDOTP = 1, S_IPROD = 2, S_GAMMA1 = 3, S_BIPROD = 4, S_BETA_NUM = 5, S_BETA_DEN = 6, S_TMP = 7, S_SRC = 8, S_OUTFLUX = 9,
       S_DIAG = 10, S_DT = 20, S_ALPHA_PREV = 30, S_ERR = 31, S_NSCAL = 32 };

static int fail(impec_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}
#define CUDA_TRY(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, IMPEC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define TRY(call) do { int rc_ = (call); if (rc_ != IMPEC_OK) return rc_; } while (0)

static inline dim3 grid2d(int nx, int ny) { return dim3((unsigned)((nx + BX - 1) / BX), (unsigned)((ny + BY - 1) / BY)); }
static const dim3 kBlock(BX, BY);
// k_pcg_matvec_march: a CTA marches down `rows` rows of a 256-column strip (64 rows; 32 on short slabs so that the grid
// still fills the 148 SMs several times over)
// (IMPEC_MARCH_ROWS: test hook -- short chunks so that small grids have interior chunks and exercise the overlapped halo exchange)
static inline int march_rows_override() { static const int v = [] { const char *e = getenv("IMPEC_MARCH_ROWS"); const int r = e ? atoi(e) : 0; return (r >= 4 && r % 4 == 0) ? r : 0; }(); return v; }
static inline int matvec_march_rows(const Grid &g) { if (march_rows_override()) return march_rows_override(); return g.N1 > 4096 ? 64 : 32; }
static inline dim3 matvec_march_grid(const Grid &g) { const int r = matvec_march_rows(g); return dim3((unsigned)((g.N2 + 2 * MV_BLOCK - 1) / (2 * MV_BLOCK)), (unsigned)((g.N1 + r - 1) / r)); }
static inline dim3 update_xr_grid(const Grid &g) { return dim3((unsigned)((g.N2 + 2 * UX_BLOCK - 1) / (2 * UX_BLOCK)), (unsigned)((g.N1 + UX_ROWS - 1) / UX_ROWS)); }

static int dev_alloc(impec_ctx *ctx, double **ptr, size_t n) {
    void *p = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&p, n * sizeof(double)));
    CUDA_TRY(ctx, cudaMemsetAsync(p, 0, n * sizeof(double), ctx->stream));
    ctx->allocs.push_back(p);
    *ptr = (double *)p;
    return IMPEC_OK;
}

static int check_launch(impec_ctx *ctx, const char *what) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, IMPEC_ERR_CUDA, "kernel %s: %s", what, cudaGetErrorString(e));
    return IMPEC_OK;
}
#define LAUNCH(ctx, kernel, grid, block, ...) do { IMPEC_LAUNCH(kernel, grid, block, 0, (ctx)->stream, __VA_ARGS__); TRY(check_launch(ctx, #kernel)); } while (0)
#define LAUNCH_COOP_SMEM(ctx, kernel, grid, block, smem, ...) do { IMPEC_LAUNCH_COOP(kernel, grid, block, smem, (ctx)->stream, __VA_ARGS__); TRY(check_launch(ctx, #kernel)); } while (0)
#define LAUNCH_COOP(ctx, kernel, grid, block, ...) do { IMPEC_LAUNCH_COOP(kernel, grid, block, 0, (ctx)->stream, __VA_ARGS__); TRY(check_launch(ctx, #kernel)); } while (0)

// ------------------------------------------------------------------------------------------------ misc kernels
__global__ void k_deinterleave(Grid g, const double *__restrict__ aos, int H, int W, double *__restrict__ u, double *__restrict__ v, int ng) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)H * W) return;
    const int i = (int)(t / W) - ng, j = (int)(t % W) - ng;
    u[gidx(g, i, j)] = aos[2 * t];
    v[gidx(g, i, j)] = aos[2 * t + 1];
}
__global__ void k_interleave(Grid g, double *__restrict__ aos, int H, int W, const double *__restrict__ u, const double *__restrict__ v, int ng) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)H * W) return;
    const int i = (int)(t / W) - ng, j = (int)(t % W) - ng;
    aos[2 * t] = u[gidx(g, i, j)];
    aos[2 * t + 1] = v[gidx(g, i, j)];
}
// acc += |val|   (HyperbolicSolver::SourceSplit, src/HyperbolicSolver.cpp:201-205)
__global__ void k_accum_abs(const double *__restrict__ val, double *__restrict__ acc) { acc[0] += fabs(val[0]); }
__global__ void k_set_scalar(double *__restrict__ dst, double v) { dst[0] = v; }
// diagonal shift that represents the fine-level Dirichlet pin on the coarse levels:
// sigma_1 = 1/8 * sum of the links between the pinned cell and its siblings in its 2x2 block, sigma_{l+1} = sigma_l/4
__global__ void k_pin_sigma(Grid g, const double *__restrict__ AW, const double *__restrict__ AS, const double *__restrict__ AV,
                            int pi, int pj, int nlev, double *__restrict__ sigma) {
    const int bi = (pi >> 1) << 1, bj = (pj >> 1) << 1;
    double tot = 0.;
    for (int di = -1; di <= 1; di++)
        for (int dj = -1; dj <= 1; dj++) {
            if (!di && !dj) continue;
            const int ii = pi + di, jj = pj + dj;
            if (ii < bi || ii > bi + 1 || jj < bj || jj > bj + 1) continue;
            if (dj == 0) tot += AW[gidx(g, pi > ii ? pi : ii, pj)];
            else if (di == 0) tot += AS[gidx(g, pi, pj > jj ? pj : jj)];
            else tot += AV[gidx(g, pi > ii ? pi : ii, pj > jj ? pj : jj)];
        }
    double sg = 0.125 * tot;
    sigma[0] = 0.;
    for (int l = 1; l < nlev; l++) { sigma[l] = sg; sg *= 0.25; }
}
// The same for a pinned ROW pi (source_type horizontal): coarse column J of level 1 holds the pinned cells (pi, 2J), (pi, 2J+1); their free
// siblings are the two cells of the other row of the 2 x 2 block: one face link and one vertex link each.
__global__ void k_pin_sigma_row(Grid g, const double *__restrict__ AW, const double *__restrict__ AV, int pi, int n, double *__restrict__ sigma1) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    if (J >= n) return;
    const int iv = (pi & 1) ? pi : pi + 1;                  // row index of the links between row pi and its sibling row
    const double ta = AW[gidx(g, iv, 2 * J)] + AV[gidx(g, iv, 2 * J + 1)];
    const double tb = AV[gidx(g, iv, 2 * J + 1)] + AW[gidx(g, iv, 2 * J + 1)];
    sigma1[J] = 0.125 * ta + 0.125 * tb;
}
__global__ void k_sigma_coarsen(int n, const double *__restrict__ fine, double *__restrict__ coarse) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    if (J < n) coarse[J] = 0.25 * (fine[2 * J] + fine[2 * J + 1]);
}
// Dirichlet values into x: value for a single cell, or scale * q for a row (its own equation kappa/dA p = -q)
__global__ void k_set_pins(double *__restrict__ x, const double *__restrict__ q, long long pin, long long pin_end, double value, double scale) {
    const long long k = pin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < pin_end) x[k] = scale != 0. ? scale * q[k] : value;
}
__global__ void k_copy_interior(Grid g, const double *__restrict__ src, double *__restrict__ dst) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.N1 || j >= g.N2) return;
    dst[gidx(g, i, j)] = src[gidx(g, i, j)];
}

// ------------------------------------------------------------------------------------------------ reductions
static int comm_allreduce(impec_ctx *ctx, double *buf, int n, int op);
static int reduce_final(impec_ctx *ctx, int op, long long nparts, int count, long long stride, int scal_idx) {
    double *out = ctx->scal + scal_idx;
    if (op == RED_SUM) LAUNCH_COOP(ctx, k_reduce_final<RED_SUM>, dim3(count), dim3(1024), ctx->part, nparts, stride, out);
    else if (op == RED_MAX) LAUNCH_COOP(ctx, k_reduce_final<RED_MAX>, dim3(count), dim3(1024), ctx->part, nparts, stride, out);
    else LAUNCH_COOP(ctx, k_reduce_final<RED_MIN>, dim3(count), dim3(1024), ctx->part, nparts, stride, out);
    return comm_allreduce(ctx, out, count, op);
}
static int fetch_scalars(impec_ctx *ctx, int idx, int n) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal + idx, ctx->scal + idx, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scal + S_ERR, ctx->scal + S_ERR, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
#endif
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p && ctx->h_scal[S_ERR] != 0.) return fail(ctx, IMPEC_ERR_NCCL, "peer-memory exchange timed out waiting for a neighbouring rank");
#endif
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ communication
// One process per GPU; fine levels are slabs of rows.  Product build: NCCL point-to-point (halo rows) and tiny
// collectives on the context stream.  Emulated build: host callbacks supplied by the test harness (gloo).
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
#define NCCL_TRY(ctx, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(ctx, IMPEC_ERR_NCCL, "%s failed: %s", #call, ncclGetErrorString(r_)); } while (0)
#endif
static int comm_check(impec_ctx *ctx) {
    if (ctx->nranks == 1) return IMPEC_OK;
    if (!ctx->comm_ready) return fail(ctx, IMPEC_ERR_NCCL, "multi-rank context without a communicator (pass an NCCL id or set callbacks)");
    return IMPEC_OK;
}
static int comm_sendrecv(impec_ctx *ctx, const double *send_lo, double *recv_lo, const double *send_hi, double *recv_hi, long long n) {
    TRY(comm_check(ctx));
    if (ctx->has_cb) { if (ctx->cb.sendrecv(ctx->cb.user, send_lo, recv_lo, send_hi, recv_hi, n)) return fail(ctx, IMPEC_ERR_NCCL, "sendrecv callback failed"); return IMPEC_OK; }
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    NCCL_TRY(ctx, ncclGroupStart());
    if (send_lo) { NCCL_TRY(ctx, ncclSend(send_lo, (size_t)n, ncclDouble, ctx->rank - 1, ctx->nccl, ctx->stream)); NCCL_TRY(ctx, ncclRecv(recv_lo, (size_t)n, ncclDouble, ctx->rank - 1, ctx->nccl, ctx->stream)); }
    if (send_hi) { NCCL_TRY(ctx, ncclSend(send_hi, (size_t)n, ncclDouble, ctx->rank + 1, ctx->nccl, ctx->stream)); NCCL_TRY(ctx, ncclRecv(recv_hi, (size_t)n, ncclDouble, ctx->rank + 1, ctx->nccl, ctx->stream)); }
    NCCL_TRY(ctx, ncclGroupEnd());
    return IMPEC_OK;
#else
    return fail(ctx, IMPEC_ERR_NCCL, "no communication back end");
#endif
}
// Periodic variant (the hyperbolic self test on slabs): the first rank's lower neighbour is the last rank.  With two ranks both neighbours are
// the same peer, so the messages are told apart by direction: what leaves through `send_lo` travels DOWN and arrives in the peer's `recv_hi`.
// Host callbacks receive all four pointers and pair them the same way (peers modulo the rank count, one tag per direction).  NCCL matches
// sends and receives between a pair of ranks in the order they are issued inside the group: sends down, up; receives from above, from below.
static int comm_sendrecv_wrap(impec_ctx *ctx, const double *send_lo, double *recv_lo, const double *send_hi, double *recv_hi, long long n) {
    TRY(comm_check(ctx));
    if (ctx->has_cb) { if (ctx->cb.sendrecv(ctx->cb.user, send_lo, recv_lo, send_hi, recv_hi, n)) return fail(ctx, IMPEC_ERR_NCCL, "sendrecv callback failed"); return IMPEC_OK; }
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    const int lo = (ctx->rank + ctx->nranks - 1) % ctx->nranks, hi = (ctx->rank + 1) % ctx->nranks;
    NCCL_TRY(ctx, ncclGroupStart());
    NCCL_TRY(ctx, ncclSend(send_lo, (size_t)n, ncclDouble, lo, ctx->nccl, ctx->stream));
    NCCL_TRY(ctx, ncclSend(send_hi, (size_t)n, ncclDouble, hi, ctx->nccl, ctx->stream));
    NCCL_TRY(ctx, ncclRecv(recv_hi, (size_t)n, ncclDouble, hi, ctx->nccl, ctx->stream));
    NCCL_TRY(ctx, ncclRecv(recv_lo, (size_t)n, ncclDouble, lo, ctx->nccl, ctx->stream));
    NCCL_TRY(ctx, ncclGroupEnd());
    return IMPEC_OK;
#else
    return fail(ctx, IMPEC_ERR_NCCL, "no communication back end");
#endif
}
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
static int p2p_allreduce(impec_ctx *ctx, double *buf, int n, int op);
#endif
static int comm_allreduce(impec_ctx *ctx, double *buf, int n, int op) {
    if (ctx->nranks == 1) return IMPEC_OK;
    TRY(comm_check(ctx));
    if (ctx->has_cb) { if (ctx->cb.allreduce(ctx->cb.user, buf, n, op)) return fail(ctx, IMPEC_ERR_NCCL, "allreduce callback failed"); return IMPEC_OK; }
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p && n <= P2P_AR_MAXN) return p2p_allreduce(ctx, buf, n, op);
    NCCL_TRY(ctx, ncclAllReduce(buf, buf, (size_t)n, ncclDouble, op == RED_SUM ? ncclSum : (op == RED_MAX ? ncclMax : ncclMin), ctx->nccl, ctx->stream));
    return IMPEC_OK;
#else
    return fail(ctx, IMPEC_ERR_NCCL, "no communication back end");
#endif
}
static int comm_allgather(impec_ctx *ctx, const double *send, double *recv, long long n) {
    TRY(comm_check(ctx));
    if (ctx->has_cb) { if (ctx->cb.allgather(ctx->cb.user, send, recv, n)) return fail(ctx, IMPEC_ERR_NCCL, "allgather callback failed"); return IMPEC_OK; }
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    NCCL_TRY(ctx, ncclAllGather(send, recv, (size_t)n, ncclDouble, ctx->nccl, ctx->stream));
    return IMPEC_OK;
#else
    return fail(ctx, IMPEC_ERR_NCCL, "no communication back end");
#endif
}
static int comm_bcast(impec_ctx *ctx, double *buf, long long n, int root) {
    if (ctx->nranks == 1) return IMPEC_OK;
    TRY(comm_check(ctx));
    if (ctx->has_cb) { if (ctx->cb.bcast(ctx->cb.user, buf, n, root)) return fail(ctx, IMPEC_ERR_NCCL, "bcast callback failed"); return IMPEC_OK; }
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    NCCL_TRY(ctx, ncclBroadcast(buf, buf, (size_t)n, ncclDouble, root, ctx->nccl, ctx->stream));
    return IMPEC_OK;
#else
    return fail(ctx, IMPEC_ERR_NCCL, "no communication back end");
#endif
}
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
// ---- peer-memory transport (comm_p2p.inl)
static const size_t kP2PHaloOff = P2P_FLAG_BYTES + P2P_AR_BYTES;
static int p2p_setup(impec_ctx *ctx) {
    const int R = ctx->nranks;
    if (R < 2 || !ctx->comm_ready) return IMPEC_OK;
    const char *e = getenv("IMPEC_P2P");
    int ok = (e && atoi(e) == 0) ? 0 : 1;
    if (R > P2P_MAXR) ok = 0;
    // mailbox: flags | all-reduce records | 2 sides x 2 parities of (2 fields x IMPEC_GH rows of the fine level)
    ctx->p2p_slot_bytes = (size_t)2 * IMPEC_GH * ctx->g.ld * sizeof(double);
    // ... | 2 parities of an all-gather window holding the interior rows of the first replicated level
    ctx->p2p_win_bytes = 0;
    if (ctx->Ld > 0 && ctx->Ld < (int)ctx->L.size()) { const Grid &gl = ctx->L[ctx->Ld].gl; ctx->p2p_win_bytes = (size_t)R * gl.N1 * gl.ld * sizeof(double); }
    const size_t bytes = kP2PHaloOff + 4 * ctx->p2p_slot_bytes + 2 * ctx->p2p_win_bytes;
    void *m = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&m, bytes));
    CUDA_TRY(ctx, cudaMemsetAsync(m, 0, bytes, ctx->stream));
    ctx->allocs.push_back(m); ctx->mail = (unsigned char *)m;
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, m) != cudaSuccess) { cudaGetLastError(); ok = 0; memset(&mine, 0, sizeof(mine)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    double *hb = nullptr;
    TRY(dev_alloc(ctx, &hb, (size_t)R * 8));
    CUDA_TRY(ctx, cudaMemcpyAsync((char *)hb + (size_t)ctx->rank * 64, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, ncclAllGather((char *)hb + (size_t)ctx->rank * 64, hb, 64, ncclChar, ctx->nccl, ctx->stream));
    std::vector<cudaIpcMemHandle_t> all(R);
    CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), hb, (size_t)R * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < R && ok; r++) {
        if (r == ctx->rank) { ctx->peer_mail[r] = ctx->mail; continue; }
        void *pm = nullptr;
        if (cudaIpcOpenMemHandle(&pm, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        ctx->peer_mail[r] = (unsigned char *)pm;
    }
    // every rank must take the same path
    double flag = ok;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal + S_TMP, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, ncclAllReduce(ctx->scal + S_TMP, ctx->scal + S_TMP, 1, ncclDouble, ncclMin, ctx->nccl, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(&flag, ctx->scal + S_TMP, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->p2p = flag > 0.5;
    { const char *eh = getenv("IMPEC_P2P_HALO"); ctx->p2p_halo = ctx->p2p && (eh ? atoi(eh) != 0 : true); }
    { const char *eg = getenv("IMPEC_P2P_GATHER"); ctx->p2p_gather = ctx->p2p && ctx->p2p_win_bytes > 0 && (eg ? atoi(eg) != 0 : true); }
    if (!ctx->p2p)
        for (int r = 0; r < R; r++) { if (r != ctx->rank && ctx->peer_mail[r]) cudaIpcCloseMemHandle(ctx->peer_mail[r]); ctx->peer_mail[r] = nullptr; }
    return IMPEC_OK;
}
static int p2p_exchange(impec_ctx *ctx, double *a, double *b, const Grid &g, int n) {
    HaloArgs h;
    h.fld[0] = a; h.fld[1] = b; h.nfld = b ? 2 : 1;
    h.cnt = (long long)n * g.ld;
    const long long ld = g.ld;
    h.send_lo = (long long)IMPEC_GH * ld; h.recv_lo = (long long)(IMPEC_GH - n) * ld;
    h.send_hi = (long long)(g.N1 - n + IMPEC_GH) * ld; h.recv_hi = (long long)(g.N1 + IMPEC_GH) * ld;
    h.epoch = ++ctx->halo_epoch;
    const size_t par = (size_t)(h.epoch & 1);
    unsigned char *me = ctx->mail;
    unsigned char *prev = g.wall_lo ? nullptr : ctx->peer_mail[ctx->rank - 1], *next = g.wall_hi ? nullptr : ctx->peer_mail[ctx->rank + 1];
    // side 0 of a mailbox receives from the owner's previous rank, side 1 from its next rank
    h.rem_lo = prev ? (double *)(prev + kP2PHaloOff + (2 + par) * ctx->p2p_slot_bytes) : nullptr;
    h.rem_lo_flag = prev ? (unsigned long long *)prev + 64 : nullptr;
    h.rem_hi = next ? (double *)(next + kP2PHaloOff + (0 + par) * ctx->p2p_slot_bytes) : nullptr;
    h.rem_hi_flag = next ? (unsigned long long *)next : nullptr;
    h.my_lo = (const double *)(me + kP2PHaloOff + (0 + par) * ctx->p2p_slot_bytes); h.my_lo_flag = (const unsigned long long *)me;
    h.my_hi = (const double *)(me + kP2PHaloOff + (2 + par) * ctx->p2p_slot_bytes); h.my_hi_flag = (const unsigned long long *)me + 64;
    h.err = ctx->scal + S_ERR;
    LAUNCH(ctx, k_halo_p2p, dim3(P2P_NCTA), dim3(P2P_THREADS), h);
    return IMPEC_OK;
}
// all-gather of `n` doubles per rank (contiguous rows of a slab) into dst_rank0 + rank * n on every rank
static int p2p_allgather(impec_ctx *ctx, const double *send, double *recv, long long n) {
    GatherArgs a;
    a.src = send; a.cnt = n; a.dst = recv; a.rank = ctx->rank; a.nranks = ctx->nranks;
    a.epoch = ++ctx->ag_epoch;
    const size_t off = kP2PHaloOff + 4 * ctx->p2p_slot_bytes + (size_t)(a.epoch & 1) * ctx->p2p_win_bytes;
    for (int r = 0; r < P2P_MAXR; r++) {
        a.rem_win[r] = r < ctx->nranks ? (double *)(ctx->peer_mail[r] + off) : nullptr;
        a.rem_flag[r] = r < ctx->nranks ? (unsigned long long *)ctx->peer_mail[r] : nullptr;
    }
    a.err = ctx->scal + S_ERR;
    LAUNCH(ctx, k_allgather_p2p, dim3(P2P_AG_NCTA), dim3(P2P_THREADS), a);
    return IMPEC_OK;
}
static int p2p_allreduce(impec_ctx *ctx, double *buf, int n, int op) {
    ArArgs a;
    a.buf = buf; a.n = n; a.op = op; a.rank = ctx->rank; a.nranks = ctx->nranks;
    for (int r = 0; r < P2P_MAXR; r++) a.mail[r] = r < ctx->nranks ? (double *)(ctx->peer_mail[r] + P2P_FLAG_BYTES) : nullptr;
    a.epoch = ++ctx->ar_epoch;
    a.err = ctx->scal + S_ERR;
    LAUNCH(ctx, k_allreduce_p2p, dim3(1), dim3(32), a);
    return IMPEC_OK;
}
#endif

// Halo rows of a slab field: my first n rows go to the previous rank's rows [N1, N1+n), my last n rows to the next
// rank's rows [-n, 0).  Works unchanged for face/vertex arrays (their row N1 duplicates the neighbour's row 0).
static int exchange_rows(impec_ctx *ctx, double *a, const Grid &g, int n) {
    if (ctx->nranks == 1 || (g.wall_lo && g.wall_hi)) return IMPEC_OK;
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p_halo && !ctx->has_cb && (size_t)n * g.ld * sizeof(double) <= ctx->p2p_slot_bytes) return p2p_exchange(ctx, a, nullptr, g, n);
#endif
    const size_t ld = (size_t)g.ld;
    const bool prev = !g.wall_lo, next = !g.wall_hi;
    return comm_sendrecv(ctx, prev ? a + (size_t)IMPEC_GH * ld : nullptr, prev ? a + (size_t)(IMPEC_GH - n) * ld : nullptr,
                         next ? a + (size_t)(g.N1 - n + IMPEC_GH) * ld : nullptr, next ? a + (size_t)(g.N1 + IMPEC_GH) * ld : nullptr, (long long)n * g.ld);
}
// The same with the ranks closed to a ring (periodic domain): every rank has both neighbours.  Plain NCCL send / recv (a test mode: the
// peer-memory mailboxes are laid out for walls at both ends of the rank line).
static int exchange_rows_periodic(impec_ctx *ctx, double *a, const Grid &g, int n) {
    if (ctx->nranks == 1) return IMPEC_OK;
    const size_t ld = (size_t)g.ld;
    return comm_sendrecv_wrap(ctx, a + (size_t)IMPEC_GH * ld, a + (size_t)(IMPEC_GH - n) * ld, a + (size_t)(g.N1 - n + IMPEC_GH) * ld, a + (size_t)(g.N1 + IMPEC_GH) * ld, (long long)n * g.ld);
}
// two fields in one NCCL group (one launch instead of two)
static int exchange_rows2(impec_ctx *ctx, double *a, double *b, const Grid &g, int n) {
    if (ctx->nranks == 1 || (g.wall_lo && g.wall_hi)) return IMPEC_OK;
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p_halo && !ctx->has_cb && b && (size_t)2 * n * g.ld * sizeof(double) <= ctx->p2p_slot_bytes) return p2p_exchange(ctx, a, b, g, n);
    if (!ctx->has_cb && b) {
        TRY(comm_check(ctx));
        const size_t ld = (size_t)g.ld; const size_t cnt = (size_t)n * ld;
        NCCL_TRY(ctx, ncclGroupStart());
        for (double *f : {a, b}) {
            if (!g.wall_lo) { NCCL_TRY(ctx, ncclSend(f + (size_t)IMPEC_GH * ld, cnt, ncclDouble, ctx->rank - 1, ctx->nccl, ctx->stream)); NCCL_TRY(ctx, ncclRecv(f + (size_t)(IMPEC_GH - n) * ld, cnt, ncclDouble, ctx->rank - 1, ctx->nccl, ctx->stream)); }
            if (!g.wall_hi) { NCCL_TRY(ctx, ncclSend(f + (size_t)(g.N1 - n + IMPEC_GH) * ld, cnt, ncclDouble, ctx->rank + 1, ctx->nccl, ctx->stream)); NCCL_TRY(ctx, ncclRecv(f + (size_t)(g.N1 + IMPEC_GH) * ld, cnt, ncclDouble, ctx->rank + 1, ctx->nccl, ctx->stream)); }
        }
        NCCL_TRY(ctx, ncclGroupEnd());
        return IMPEC_OK;
    }
#endif
    TRY(exchange_rows(ctx, a, g, n));
    if (b) TRY(exchange_rows(ctx, b, g, n));
    return IMPEC_OK;
}
// ---- halo exchange overlapped with interior compute
// A marching kernel works on chunks of rows; only the first and the last chunk of a slab read halo rows.  `xchg` issues the exchange
// of exactly the rows the kernel reads, `launch(y0, ys, ny)` launches the kernel on the ny chunks y0, y0 + ys, ...  Overlapped form:
//     main stream :  [producer] -- fork --> interior chunks (1 .. ny-2) ----------------------- join --> [next kernel]
//     side stream :               fork --> exchange --> chunk 0 and chunk ny-1 (the halo readers) --> join
// so the NVLink round trip of the exchange (17-20 us) and the two edge chunks hide behind the interior chunks.  The side stream has
// the higher priority: its CTAs are placed as soon as interior CTAs retire.  Same kernels, same arithmetic: bit-identical.
template <class X, class K>
static int halo_overlapped(impec_ctx *ctx, bool dist, int ny, int rows_per_chunk, int N1, X xchg, K launch) {
    // (the last chunk must be at least 4 rows tall, or the chunk before it would read halo rows as well)
    const bool ov = dist && ctx->overlap && ctx->nranks > 1 && ny >= 4 && rows_per_chunk >= 4 && N1 - (ny - 1) * rows_per_chunk >= 4;
    if (!ov) { if (dist) TRY(xchg()); return launch(0, 1, ny); }
    ctx->overlapped++;
#ifndef IMPEC_EMU
    if (!ctx->has_cb && ctx->side) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
        cudaStream_t main = ctx->stream;
        ctx->stream = ctx->side;
        int rc = xchg();
        if (rc == IMPEC_OK) rc = launch(0, ny - 1, 2);
        ctx->stream = main;
        if (rc != IMPEC_OK) return rc;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->side));
        TRY(launch(1, 1, ny - 2));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        return IMPEC_OK;
    }
#endif
    // emulated build / host-callback transport: the same three pieces, in order
    TRY(xchg());
    TRY(launch(0, ny - 1, 2));
    return launch(1, 1, ny - 2);
}
static LevelDev chunked(LevelDev D, int y0, int ys, int pfd) { D.cy0 = y0; D.cys = ys; D.pfd = pfd; return D; }

extern "C" int impec_ctx_set_comm_callbacks(impec_ctx *ctx, const impec_comm_callbacks *cb) {
    if (!cb || !cb->sendrecv || !cb->allreduce || !cb->allgather || !cb->bcast) return fail(ctx, IMPEC_ERR_ARG, "incomplete callback table");
#ifndef IMPEC_EMU
    return fail(ctx, IMPEC_ERR_UNSUPPORTED, "host callbacks are for the emulated build; the CUDA library uses NCCL");
#else
    ctx->cb = *cb; ctx->has_cb = true; ctx->comm_ready = true;
    return IMPEC_OK;
#endif
}
extern "C" int impec_local_rows(const impec_ctx *ctx, int *row0, int *nrows) {
    if (row0) *row0 = ctx->g.i_off; if (nrows) *nrows = ctx->g.N1;
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ setup
static void fill_dev_params(impec_ctx *ctx) {
    const impec_params &P = ctx->P;
    DevParams &D = ctx->D;
    D.dx = (P.x_ub - P.x_lb) / P.N1; D.dy = (P.y_ub - P.y_lb) / P.N2;       // src/SimParams.cpp:452-453
    D.x_lb = P.x_lb; D.y_lb = P.y_lb;
    D.mu_res = P.mu_res; D.mq = pow(P.mu_res / P.mu_inj, 0.25);              // src/EllipticSolver.cpp:126
    D.nu = P.nu; D.porosity = P.porosity; D.inj_fluid_conc = P.inj_fluid_conc;
    D.alpha_l = P.alpha_l; D.alpha_t = P.alpha_t;
    const double alpha = atan(D.dx / D.dy), theta = atan(D.dy / D.dx);       // src/EllipticSolver.cpp:72-93
    D.dxi = D.dx / cos(theta); D.deta = D.dy / cos(alpha);
    const double secamt = 1. / cos(alpha - theta);
    D.secamt2 = secamt * secamt;
    D.cosa2 = cos(alpha) * cos(alpha); D.sina2 = sin(alpha) * sin(alpha);
    D.costsina = cos(theta) * sin(alpha); D.cosasint = cos(alpha) * sin(theta);
    D.cost2 = cos(theta) * cos(theta); D.sint2 = sin(theta) * sin(theta);
    D.cosa = cos(alpha); D.sint = sin(theta); D.cost = cos(theta); D.sina = sin(alpha);
    D.avgproc = P.avgproc; D.is_test = P.is_test; D.testno = P.is_test ? P.testno : 0;
}

extern "C" void impec_default_params(impec_params *P) {
    memset(P, 0, sizeof(*P));
    P->N1 = P->N2 = 64; P->g = 2;
    P->x_ub = P->y_ub = 1.0;
    P->mu_res = 41.0; P->mu_inj = 1.0; P->avgproc = IMPEC_HARMONIC;
    P->nu = 2.0 / 3.0; P->stype = IMPEC_PARALLEL;
    P->porosity = 1.0; P->cfl = 0.5; P->inj_fluid_conc = 1.0;
    P->alpha_l = P->alpha_t = 1.0; P->dt_diff_limit = 0.9; P->ptype = IMPEC_CRANKNICOLSON;
    P->eps = 1e-14; P->para_eps = 1e-14;
    P->hyper_method = IMPEC_MUSCL; P->dim_splitting = IMPEC_SYMMETRIC_PARALLEL;
    P->max_iter = 10000; P->mg_max_levels = 0; P->mg_nmin = 4; P->mg_ncoarse = 8; P->mg_w_from = 2; P->mg_w_to = 4;
}

extern "C" const char *impec_version(void) { return "impec2d_b200 0.1 (sm_100a)"; }
extern "C" int impec_is_emulated(void) {
#ifdef IMPEC_EMU
    return 1;
#else
    return 0;
#endif
}
// Host placement (include/impec2d_b200.h): the CPUs local to a device, from the kernel's view of its PCI function.
// "0-55,112-167" -> cpu set; false on a malformed list
static bool parse_cpulist(const char *s, cpu_set_t *set) {
    CPU_ZERO(set);
    const char *p = s;
    while (*p && *p != '\n') {
        char *e;
        const long a = strtol(p, &e, 10);
        if (e == p || a < 0) return false;
        long b = a;
        p = e;
        if (*p == '-') { b = strtol(p + 1, &e, 10); if (e == p + 1 || b < a) return false; p = e; }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) CPU_SET((int)c, set);
        if (*p == ',') p++;
        else if (*p && *p != '\n') return false;
    }
    return true;
}
extern "C" int impec_bind_host_to_device(int device, char *cpulist_out, int len) {
    if (cpulist_out && len > 0) cpulist_out[0] = 0;
    char line[4096] = {0};
#ifdef IMPEC_EMU
    (void)device;
    const char *fake = getenv("IMPEC_EMU_LOCAL_CPULIST");          // test hook of the emulated build: there is no PCI device to ask
    if (!fake) return IMPEC_ERR_UNSUPPORTED;
    snprintf(line, sizeof line, "%s", fake);
#else
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return IMPEC_ERR_CUDA; }
    for (char *c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');        // sysfs spells the bus id in lower case
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return IMPEC_ERR_UNSUPPORTED;
    const bool got = fgets(line, (int)sizeof line, f) != nullptr;
    fclose(f);
    if (!got) return IMPEC_ERR_UNSUPPORTED;
#endif
    cpu_set_t want, cur, both;
    if (!parse_cpulist(line, &want) || CPU_COUNT(&want) == 0) return IMPEC_ERR_UNSUPPORTED;
    if (sched_getaffinity(0, sizeof cur, &cur) != 0) return IMPEC_ERR_UNSUPPORTED;
    CPU_AND(&both, &want, &cur);
    const int n = CPU_COUNT(&both);
    if (n == 0) return IMPEC_ERR_UNSUPPORTED;                     // e.g. a container cpuset on the other socket: leave the mask alone
    if (sched_setaffinity(0, sizeof both, &both) != 0) return IMPEC_ERR_UNSUPPORTED;
    if (cpulist_out && len > 0) { snprintf(cpulist_out, (size_t)len, "%s", line); for (char *c = cpulist_out; *c; ++c) if (*c == '\n') *c = 0; }
    return n;
}
extern "C" int impec_nccl_unique_id(char id_out[128]) {
    memset(id_out, 0, 128);
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId does not fit the ABI buffer");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return IMPEC_ERR_NCCL;
    memcpy(id_out, &id, sizeof(id));
    return IMPEC_OK;
#else
    return IMPEC_ERR_UNSUPPORTED;
#endif
}

extern "C" const char *impec_last_error(const impec_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" void impec_ctx_destroy(impec_ctx *ctx) {
    if (!ctx) return;
    if (getenv("IMPEC_VERBOSE")) fprintf(stderr, "impec2d_b200: %lld kernel launches, %lld of them marching multigrid kernels, %lld halo exchanges overlapped with interior chunks\n", ctx->launches, ctx->march_launches, ctx->overlapped);
    cudaStreamSynchronize(ctx->stream);
    for (void *p : ctx->allocs) cudaFree(p);
    for (void *p : {(void *)ctx->well_i, (void *)ctx->well_j, (void *)ctx->well_q, (void *)ctx->well_md}) if (p) cudaFree(p);
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p) for (int r = 0; r < ctx->nranks; r++) if (r != ctx->rank && ctx->peer_mail[r]) cudaIpcCloseMemHandle(ctx->peer_mail[r]);
    if (ctx->nccl) ncclCommDestroy(ctx->nccl);
#endif
#ifndef IMPEC_EMU
    if (ctx->tail_exec) cudaGraphExecDestroy(ctx->tail_exec);
    if (ctx->tail_graph) cudaGraphDestroy(ctx->tail_graph);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
#endif
#ifdef IMPEC_EMU
    if (ctx->filecomm) { ctx->filecomm->finish(); delete ctx->filecomm; }
#endif
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    for (void *p : {(void *)ctx->hw_i, (void *)ctx->hw_j, (void *)ctx->hw_q}) if (p) cudaFreeHost(p);
#ifndef IMPEC_EMU
    if (ctx->ev_wells) cudaEventDestroy(ctx->ev_wells);
#endif
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int impec_ctx_create(const impec_params *P, int device, int rank, int nranks, const char *nccl_unique_id, void *cuda_stream, impec_ctx **out) {
    if (!P || !out) return IMPEC_ERR_ARG;
    *out = nullptr;
    impec_ctx *ctx = new impec_ctx();
    *out = ctx;                                   // returned even on failure so that the error text can be read
    ctx->P = *P; ctx->device = device; ctx->rank = rank; ctx->nranks = nranks;
    // Cycle shape by rank count.  The multigrid levels below 4 M cells are replicated on every rank, so their cost does not shrink with the
    // number of GPUs; once a slab is small enough that this tail is a third of the cycle, entering it 4 x 1 instead of 4 x 2 times per cycle
    // (W range 2..3 instead of the default 2..4) more than pays for the extra PCG iterations: 16384^2 on 8 B200, 23 iterations at 57.8 ms per
    // step against 20 at 60.5 (profiles/r2ad_*); on 1 / 2 / 4 GPUs (and for larger slabs) the default range stays the faster one
    // (profiles/r2x_cycle_shapes.txt).  Only the default range is touched; impec_set_option("mg_w_to", ..) overrides, IMPEC_W_RANGE_AUTO=0 disables.
    {
        const char *e = getenv("IMPEC_W_RANGE_AUTO");
        const double cells_per_rank = (double)P->N1 * P->N2 / (nranks > 0 ? nranks : 1);
        if ((!e || atoi(e) != 0) && nranks >= 8 && P->mg_w_from == 2 && P->mg_w_to == 4 && cells_per_rank >= 1.0e7 && cells_per_rank <= 5.0e7) ctx->P.mg_w_to = 3;
    }
    if (P->g != 2) return fail(ctx, IMPEC_ERR_ARG, "num_ghostcells must be 2 (got %d)", P->g);
    if (P->N1 < 4 || P->N2 < 4) return fail(ctx, IMPEC_ERR_ARG, "grid too small: %d x %d", P->N1, P->N2);
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, IMPEC_ERR_ARG, "bad rank %d of %d", rank, nranks);
    if (nranks > 1) {
        if (P->N1 % (2 * nranks) != 0 || P->N1 / nranks < 8) return fail(ctx, IMPEC_ERR_ARG, "x_cells = %d cannot be cut into %d slabs of an even number (>= 8) of rows", P->N1, nranks);
    }
    CUDA_TRY(ctx, cudaSetDevice(device));
    if (cuda_stream) ctx->stream = (cudaStream_t)cuda_stream;
    else { CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    { const char *e = getenv("IMPEC_HALO_OVERLAP"); if (e) ctx->overlap = atoi(e) != 0; }
    { const char *e = getenv("IMPEC_DEFER_X"); if (e) ctx->defer_x = atoi(e) != 0; }
    { const char *e = getenv("IMPEC_SWEEP_PREFETCH"); if (e) ctx->pf_sweep = atoi(e); }
    { const char *e = getenv("IMPEC_SWEEP_Y_PAIRS"); if (e) ctx->sweep_y_pairs = atoi(e) != 0; }
    { const char *e = getenv("IMPEC_L2_PREFETCH");
      if (e) { int v[4]; const int n = sscanf(e, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]); for (int a = n; a < 4 && n > 0; a++) v[a] = v[n - 1];
               if (n > 0) { ctx->pf_pre = v[0]; ctx->pf_post = v[1]; ctx->pf_mv = v[2]; ctx->pf_sgs = v[3]; } } }
#ifndef IMPEC_EMU
    if (nranks > 1) {
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
#define SMEM_ATTR(kern) CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_W * FT_H * (int)sizeof(double)))
#define SMEM_ATTR3(MODE, F32) SMEM_ATTR((k_mg_fused<MODE, F32, 64>)); SMEM_ATTR((k_mg_fused<MODE, F32, 32>)); SMEM_ATTR((k_mg_fused<MODE, F32, 16>))
    SMEM_ATTR3(FUSED_PLAIN, true); SMEM_ATTR3(FUSED_PLAIN, false);
    SMEM_ATTR3(FUSED_PRE, true); SMEM_ATTR3(FUSED_PRE, false);
    SMEM_ATTR3(FUSED_POST, true); SMEM_ATTR3(FUSED_POST, false);
    SMEM_ATTR3(FUSED_MID, true); SMEM_ATTR3(FUSED_MID, false);
    SMEM_ATTR(k_mg_coarse_solve<true>); SMEM_ATTR(k_mg_coarse_solve<false>);
#undef SMEM_ATTR3
#undef SMEM_ATTR
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_pcg_matvec<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_W * FT_H * (int)sizeof(double)));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_pcg_matvec<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_W * FT_H * (int)sizeof(double)));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_mg_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_mg_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
#endif
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (nranks > 1 && nccl_unique_id) {
        ncclUniqueId id; memcpy(&id, nccl_unique_id, sizeof(id));
        NCCL_TRY(ctx, ncclCommInitRank(&ctx->nccl, nranks, id, rank));
        ctx->comm_ready = true;
    }
#else
    (void)nccl_unique_id;
#endif
#ifdef IMPEC_EMU
    if (nranks > 1 && getenv("IMPEC_EMU_COMM_DIR")) {        // emulated ranks as plain processes (the C++ driver on slabs, tests/test_host_layer.py)
        ctx->filecomm = new EmuFileComm();
        static int generation = 0;
        ctx->filecomm->dir = getenv("IMPEC_EMU_COMM_DIR"); ctx->filecomm->rank = rank; ctx->filecomm->nranks = nranks; ctx->filecomm->gen = generation++;
        ctx->cb.user = ctx->filecomm; ctx->cb.sendrecv = EmuFileComm::sendrecv; ctx->cb.allreduce = EmuFileComm::allreduce;
        ctx->cb.allgather = EmuFileComm::allgather; ctx->cb.bcast = EmuFileComm::bcast;
        ctx->has_cb = true; ctx->comm_ready = true;
    }
#endif
    fill_dev_params(ctx);
    ctx->N1g = P->N1;
    ctx->g = make_slab(P->N1, P->N2, rank, nranks);
    ctx->D.x_lb = P->x_lb + ctx->g.i_off * ctx->D.dx;       // device-side coordinates (the K(x,y) fields of the manufactured tests) count rows from the slab's first row
    const size_t n = grid_elems(ctx->g);
    double **fields[] = {&ctx->c, &ctx->p, &ctx->q, &ctx->u, &ctx->v, &ctx->MU, &ctx->KV11, &ctx->r, &ctx->z, &ctx->d, &ctx->s, &ctx->rhs, &ctx->t0, &ctx->d2buf};
    for (double **f : fields) TRY(dev_alloc(ctx, f, n));
    TRY(dev_alloc(ctx, &ctx->ta, 2 * n)); ctx->tb = ctx->ta + n;
    if (P->is_test) TRY(dev_alloc(ctx, &ctx->KV22, n)); else ctx->KV22 = ctx->KV11;
    // multigrid hierarchy
    int nl = 1, a = P->N1, b = P->N2;
    const int nmin = P->mg_nmin > 1 ? P->mg_nmin : 2;
    while ((a % 2 == 0) && (b % 2 == 0) && std::min(a, b) > nmin && (P->mg_max_levels <= 0 || nl < P->mg_max_levels)) { a /= 2; b /= 2; nl++; }
    ctx->L.resize(nl);
    // distributed levels: as long as every slab keeps an even number (>= 8) of rows; the coarsest level is always replicated
    ctx->Ld = 0;
    // Small levels are latency-bound: three halo exchanges per visit cost more than the slab saves, so levels below
    // IMPEC_DIST_MIN_CELLS cells (default 4M; 0 in the parity tests) are replicated as well.
    if (nranks > 1) {
        const char *e = getenv("IMPEC_DIST_MIN_CELLS");
        const double min_cells = e ? atof(e) : 4.0e6;
        int rows = P->N1, cols = P->N2;
        // Level 0 is ALWAYS distributed on a multi-rank context (its vectors alias the slab-sized PCG vectors); min_cells applies
        // from level 1 on.
        while (ctx->Ld < nl - 1 && rows % (2 * nranks) == 0 && rows / nranks >= 8 && (ctx->Ld == 0 || (double)rows * cols >= min_cells)) { ctx->Ld++; rows /= 2; cols /= 2; }
        if (ctx->Ld < 1) return fail(ctx, IMPEC_ERR_ARG, "a %d x %d grid on %d ranks leaves no coarser level to replicate (multigrid needs >= 2 levels on slabs)", P->N1, P->N2, nranks);
    }
    for (int l = 0, n1 = P->N1, n2 = P->N2; l < nl; l++, n1 /= 2, n2 /= 2) {
        Level &L = ctx->L[l];
        L.dist = l < ctx->Ld;
        L.g = L.dist ? make_slab(n1, n2, rank, nranks) : make_grid(n1, n2);
        if (l == ctx->Ld && nranks > 1) {
            L.gl = make_slab(n1, n2, rank, nranks);
            const size_t ns = grid_elems(L.gl);
            TRY(dev_alloc(ctx, &L.AWl, ns)); TRY(dev_alloc(ctx, &L.ASl, ns)); TRY(dev_alloc(ctx, &L.AVl, ns)); TRY(dev_alloc(ctx, &L.fl, ns));
        } else L.gl = L.g;
        {   // tile height: big levels amortise the halo best with 64-row tiles; small levels are a few tiles only, and
            // the duration of a launch is the time one CTA needs for its tile, so they get more, shorter tiles
            const char *e32 = getenv("IMPEC_TH32_BELOW"), *e16 = getenv("IMPEC_TH16_BELOW");
            const double c32 = e32 ? atof(e32) : 6.0e6, c16 = e16 ? atof(e16) : 6.0e6, cells = (double)L.g.N1 * L.g.N2;
            L.th = cells < c16 ? 16 : (cells < c32 ? 32 : 64);
            // big levels: register-march kernels.  Chunk height: as tall as possible (3 + 3 rows of every chunk are
            // redundant) while the grid still fills the 148 SMs (3 CTAs each) about three times over.
            const char *em = getenv("IMPEC_MARCH_MIN_CELLS");
            const double cm = em ? atof(em) : 8.0e6;
            L.march_rows = 0;
            if (cells >= cm && L.g.N1 % 2 == 0 && L.g.N2 % 2 == 0 && l < nl - 1) {
                const int ctas_x = ((L.g.N2 + MGP_OWN - 1) / MGP_OWN + MG_BLOCK / 32 - 1) / (MG_BLOCK / 32);
                int ch = 256;
                while (ch > 32 && (long long)ctas_x * ((L.g.N1 + ch - 1) / ch) < 1200) ch /= 2;
                L.march_rows = march_rows_override() ? march_rows_override() : ch;
            }
        }
        const size_t nn = grid_elems(L.g);
        TRY(dev_alloc(ctx, &L.AW, nn)); TRY(dev_alloc(ctx, &L.AS, nn)); TRY(dev_alloc(ctx, &L.AV, nn)); TRY(dev_alloc(ctx, &L.DINV, nn));
        { double *f32 = nullptr; TRY(dev_alloc(ctx, &f32, 2 * nn + 8)); L.AWf = (float *)f32; L.ASf = L.AWf + nn; L.AVf = L.ASf + nn; L.DINVf = L.AVf + nn; }
        if (l == 0) { L.x = ctx->z; L.f = ctx->r; L.t = ctx->t0; L.r = ctx->s; }
        else { TRY(dev_alloc(ctx, &L.x, nn)); TRY(dev_alloc(ctx, &L.f, nn)); TRY(dev_alloc(ctx, &L.t, nn)); TRY(dev_alloc(ctx, &L.r, nn)); }
    }
    for (int l = std::max(ctx->Ld, 1); l < nl; l++)
        if ((double)ctx->L[l].g.N1 * ctx->L[l].g.N2 <= 2097152.0 && l < nl - 1) { ctx->tail_level = l; break; }
    // Dirichlet cells of the pressure problem (src/MGSolver.cpp:109-140): cell (1,1), or the whole row i = 1 for source_type horizontal.
    // On the coarse levels they become a diagonal shift of the coarse cells that contain them (k_pin_sigma / k_pin_sigma_row).
    {
        const bool row = P->stype == IMPEC_HORIZONTAL;
        int pi = 1, pj = row ? 0 : 1;
        ctx->L[0].pin = (rank == 0) ? (long long)gidx(ctx->g, pi, pj) : -1;
        ctx->L[0].pin_end = (rank == 0) ? ctx->L[0].pin + (row ? P->N2 : 1) : -1;
        long long off = 1;                                            // (slot 0: level 0 has no shift)
        for (int l = 1; l < nl; l++) {
            pi >>= 1; pj >>= 1;
            Level &L = ctx->L[l];
            const int cnt = row ? L.g.N2 : 1;
            L.sig_idx = (rank == 0 || !L.dist) ? (long long)gidx(L.g, pi, pj) : -1;
            L.sig_end = L.sig_idx >= 0 ? L.sig_idx + cnt : -1;
            L.sig_off = off; off += cnt;
        }
        ctx->sigma_len = off;
        TRY(dev_alloc(ctx, &ctx->sigma_dev, (size_t)off + 1));
    }
    // reductions
    const dim3 gg = grid2d(P->N2 + 2, ctx->g.N1 + 2);
    ctx->part_cap = (long long)gg.x * gg.y;
    TRY(dev_alloc(ctx, &ctx->part, (size_t)ctx->part_cap * 4));
    TRY(dev_alloc(ctx, &ctx->scal, S_NSCAL));
    CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->h_scal, sizeof(double) * S_NSCAL));
    memset(ctx->h_scal, 0, sizeof(double) * S_NSCAL);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    TRY(p2p_setup(ctx));
#endif
    return IMPEC_OK;
}

extern "C" const char *impec_comm_backend(const impec_ctx *ctx) {
    if (!ctx || ctx->nranks == 1) return "single";
    if (ctx->has_cb) return "callbacks";
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p) return ctx->p2p_halo ? (ctx->p2p_gather ? "nvlink-p2p (halo rows, scalar all-reduce, coarse right-hand-side all-gather) + nccl (bootstrap, coefficient all-gather)" : "nvlink-p2p+nccl") : "nccl+p2p-allreduce";
    if (ctx->comm_ready) return "nccl";
#endif
    return "none";
}

extern "C" int impec_synchronize(impec_ctx *ctx) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); return IMPEC_OK; }
extern "C" int impec_mg_levels(const impec_ctx *ctx) { return (int)ctx->L.size(); }
extern "C" int impec_mg_w_range(const impec_ctx *ctx, int *from, int *to) { if (!ctx || !from || !to) return IMPEC_ERR_ARG; *from = ctx->P.mg_w_from; *to = ctx->P.mg_w_to; return IMPEC_OK; }
extern "C" long long impec_kernel_launches(const impec_ctx *ctx) { return ctx->launches; }
extern "C" int impec_last_parabolic_info(const impec_ctx *ctx, impec_solve_info *info) { if (!ctx || !info) return IMPEC_ERR_ARG; *info = ctx->last_para; return IMPEC_OK; }
extern "C" long long impec_overlapped_exchanges(const impec_ctx *ctx) { return ctx->overlapped; }

extern "C" int impec_reset_state(impec_ctx *ctx) {
    ctx->prev_first_sweep = 1;
    LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->scal + S_OUTFLUX, 0.0);
    return IMPEC_OK;
}
extern "C" int impec_set_option(impec_ctx *ctx, const char *key, double value) {
    const std::string k(key ? key : "");
    if (k == "warm_start") { ctx->warm_start = value != 0.0; return IMPEC_OK; }
    if (k == "mg_fused") { ctx->fused = value != 0.0; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "stagnation_guard") { ctx->stagnation_guard = value != 0.0; return IMPEC_OK; }
    if (k == "mg_links_fp32") { ctx->links_f32 = value != 0.0; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "fast_r0") { ctx->fast_r0 = value != 0.0; return IMPEC_OK; }
    if (k == "sgs_march") { ctx->sgs_march = value != 0.0; return IMPEC_OK; }
    if (k == "halo_overlap") { ctx->overlap = value != 0.0; return IMPEC_OK; }
    if (k == "defer_x") { ctx->defer_x = value != 0.0; return IMPEC_OK; }
    if (k == "sweep_y_pairs") { ctx->sweep_y_pairs = value != 0.0; return IMPEC_OK; }
    if (k == "sweep_prefetch") { ctx->pf_sweep = (int)value; return IMPEC_OK; }
    if (k == "l2_prefetch") { ctx->pf_pre = ctx->pf_post = ctx->pf_mv = ctx->pf_sgs = (int)value; return IMPEC_OK; }
    if (k.rfind("mg_march_rows@", 0) == 0) {           // chunk height of the marching kernels on one level (tuning; the level must already march)
        const int l = atoi(k.c_str() + 14), r = (int)value;
        if (l < 0 || l >= (int)ctx->L.size() || ctx->L[l].march_rows <= 0 || r < 16 || r % 4) return fail(ctx, IMPEC_ERR_ARG, "mg_march_rows: level %d does not march or %d is not a multiple of 4 >= 16", l, r);
        ctx->L[l].march_rows = r; ctx->graph_ready = false; return IMPEC_OK;
    }
    if (k == "mg_march") { ctx->march = value != 0.0; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "mg_fused_mid") { ctx->fused_mid = value != 0.0; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "matvec_march") { ctx->matvec_march = value != 0.0; return IMPEC_OK; }
    if (k == "sweep_shared") { ctx->sweep_shared = value != 0.0; return IMPEC_OK; }
    if (k == "vel_march") { ctx->vel_march = value != 0.0; ctx->dt_cached = false; return IMPEC_OK; }
    if (k == "build_fused") { ctx->build_fused = value != 0.0; ctx->op_valid = false; return IMPEC_OK; }
    if (k == "mg_small") { ctx->use_small = value != 0.0; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "mg_graph") { ctx->use_graph = value != 0.0; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "mg_w_from") { ctx->P.mg_w_from = (int)value; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "mg_w_to") { ctx->P.mg_w_to = (int)value; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "mg_ncoarse") { ctx->P.mg_ncoarse = (int)value; ctx->graph_ready = false; return IMPEC_OK; }
    if (k == "max_iter") { ctx->P.max_iter = (int)value; return IMPEC_OK; }
    if (k == "profile_dt") { ctx->profile_dt = value; return IMPEC_OK; }
    if (k == "hyper_method") {                   // MUSCL / upwind is a property of the calling solver object, not of the context
        if ((int)value != IMPEC_MUSCL && (int)value != IMPEC_UPWIND) return fail(ctx, IMPEC_ERR_ARG, "hyper_method must be IMPEC_MUSCL or IMPEC_UPWIND");
        ctx->P.hyper_method = (int)value; return IMPEC_OK;
    }
    return fail(ctx, IMPEC_ERR_ARG, "unknown option '%s'", k.c_str());
}

// ------------------------------------------------------------------------------------------------ transfers
// `host` points at the slab's first ghost row: row -ng of the slab = global row i_off - ng
static int upload_scalar(impec_ctx *ctx, double *dst, const double *host) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    const size_t w = (size_t)(g.N2 + 2 * ng) * sizeof(double);
    CUDA_TRY(ctx, cudaMemcpy2DAsync(dst + gidx(g, -ng, -ng), (size_t)g.ld * sizeof(double), host, w, w, (size_t)(g.N1 + 2 * ng), cudaMemcpyHostToDevice, ctx->stream));
    ctx->bytes_h2d += w * (size_t)(g.N1 + 2 * ng);
    return IMPEC_OK;
}
static int download_scalar(impec_ctx *ctx, const double *src, double *host) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    const size_t w = (size_t)(g.N2 + 2 * ng) * sizeof(double);
    CUDA_TRY(ctx, cudaMemcpy2DAsync(host, w, src + gidx(g, -ng, -ng), (size_t)g.ld * sizeof(double), w, (size_t)(g.N1 + 2 * ng), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bytes_d2h += w * (size_t)(g.N1 + 2 * ng);
    return IMPEC_OK;
}
static size_t slab_off(const impec_ctx *ctx) { return (size_t)ctx->g.i_off * (size_t)(ctx->g.N2 + 2 * ctx->P.g); }
extern "C" int impec_upload_c(impec_ctx *ctx, const double *h) { ctx->mob_valid = ctx->op_valid = false; return upload_scalar(ctx, ctx->c, h + slab_off(ctx)); }
extern "C" int impec_upload_p(impec_ctx *ctx, const double *h) { return upload_scalar(ctx, ctx->p, h + slab_off(ctx)); }
extern "C" int impec_download_c(impec_ctx *ctx, double *h) { return download_scalar(ctx, ctx->c, h + slab_off(ctx)); }
extern "C" int impec_download_p(impec_ctx *ctx, double *h) { return download_scalar(ctx, ctx->p, h + slab_off(ctx)); }
extern "C" int impec_upload_c_rows(impec_ctx *ctx, const double *h) { ctx->mob_valid = ctx->op_valid = false; return upload_scalar(ctx, ctx->c, h); }
extern "C" int impec_download_c_rows(impec_ctx *ctx, double *h) { return download_scalar(ctx, ctx->c, h); }

// q: this rank's rows (N1loc x N2), dense on the host as the reference passes it (double const *q).  The wells are a tiny part of the
// grid (0.8 % of the cells in the shipped cfgs), so the host array is SCANNED on all host threads -- while earlier uploads on the
// stream (c) are still in flight -- and only the compact list of non-zero cells crosses PCIe; the dense device array (read by
// k_setup_rhs and the diagnostics) is rebuilt by a memset + scatter kernel.  A dense q (more than 1/16 of the cells non-zero) is copied
// as it is.
__global__ void k_scatter_wells(int n, const int *__restrict__ wi, const int *__restrict__ wj, const double *__restrict__ wq, Grid g, double *__restrict__ q) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) q[gidx(g, wi[t], wj[t])] = wq[t];
}
// Non-zero cells of rows [i0, i1) of a dense row-major array; FILL = false only counts.  Blocks of 32 values that are all +-0.0 (OR of the
// bit patterns without the sign bit) are skipped without a per-element branch: the scan runs at memory speed where q is zero, i.e. almost
// everywhere.
template <bool FILL> static long long scan_rows_nonzero(const double *q, int N2, int i0, int i1, int *oi, int *oj, double *oq) {
    long long n = 0;
    for (int i = i0; i < i1; i++) {
        const double *row = q + (size_t)i * N2;
        const unsigned long long *bits = reinterpret_cast<const unsigned long long *>(row);
        int j = 0;
        for (; j + 32 <= N2; j += 32) {
            unsigned long long a = 0;
            for (int e = 0; e < 32; e++) a |= bits[j + e];
            if ((a << 1) == 0) continue;
            for (int e = 0; e < 32; e++) { const double v = row[j + e]; if (v != 0.0) { if (FILL) { oi[n] = i; oj[n] = j + e; oq[n] = v; } n++; } }
        }
        for (; j < N2; j++) { const double v = row[j]; if (v != 0.0) { if (FILL) { oi[n] = i; oj[n] = j; oq[n] = v; } n++; } }
    }
    return n;
}
// Compact list of the non-zero cells of q in row-major order, written into the context's pinned staging buffers (hw_i / hw_j / hw_q).
// Two passes over chunks of 16 rows that the host threads pull from a shared counter (the wells sit in a few row blocks: a static
// partition leaves most threads idle while one or two collect every entry): count per chunk, prefix sum, fill in place.  Only chunks
// that hold a well are read twice.
static int scan_wells(impec_ctx *ctx, const double *q, long long *nwell) {
    const int N1 = ctx->g.N1, N2 = ctx->P.N2;
    const int R = 16, nch = (N1 + R - 1) / R;
    unsigned nt = std::thread::hardware_concurrency();
    if (ctx->nranks > 1) nt /= (unsigned)std::min(ctx->nranks, 8);          // the ranks of one node share its cores
    if (nt < 2) nt = 2; if (nt > 64) nt = 64; if ((int)nt > nch) nt = (unsigned)nch;
    std::vector<long long> off((size_t)nch + 1, 0);
    std::atomic<int> next(0);
    auto run = [&](const std::function<void(int)> &body) {
        next = 0;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) th.emplace_back([&]() { for (;;) { const int c = next.fetch_add(1); if (c >= nch) break; body(c); } });
        for (auto &x : th) x.join();
    };
    run([&](int c) { off[c] = scan_rows_nonzero<false>(q, N2, c * R, std::min(N1, (c + 1) * R), nullptr, nullptr, nullptr); });
    long long tot = 0;
    for (int c = 0; c < nch; c++) { const long long n = off[c]; off[c] = tot; tot += n; }
    off[nch] = tot;
    if (tot > 0x7fffffffLL) return fail(ctx, IMPEC_ERR_ARG, "source term with more than 2^31 non-zero cells");
#ifndef IMPEC_EMU
    // the staging buffers may still be the source of the previous upload's copies
    if (ctx->ev_wells) CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_wells));
#endif
    if (tot > ctx->hw_cap) {
        for (void *old : {(void *)ctx->hw_i, (void *)ctx->hw_j, (void *)ctx->hw_q}) if (old) cudaFreeHost(old);
        ctx->hw_i = ctx->hw_j = nullptr; ctx->hw_q = nullptr; ctx->hw_cap = 0;
        const long long cap = tot + tot / 8 + 1024;
        CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->hw_i, sizeof(int) * (size_t)cap));
        CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->hw_j, sizeof(int) * (size_t)cap));
        CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->hw_q, sizeof(double) * (size_t)cap));
        ctx->hw_cap = cap;
    }
    if (tot) run([&](int c) { if (off[c + 1] > off[c]) scan_rows_nonzero<true>(q, N2, c * R, std::min(N1, (c + 1) * R), ctx->hw_i + off[c], ctx->hw_j + off[c], ctx->hw_q + off[c]); });
    *nwell = tot;
    return IMPEC_OK;
}
// device copies of the compact list (asynchronous, from the pinned staging buffers)
static int set_wells(impec_ctx *ctx, const double *q, long long nwell) {
    const int N1 = ctx->g.N1, N2 = ctx->P.N2;
    ctx->nwell = (int)nwell;
    ctx->q_pin = (ctx->rank == 0 && N1 > 1 && N2 > 1) ? q[(size_t)1 * N2 + 1] : 0.0;
    if (ctx->nwell) {
        if (ctx->nwell > ctx->well_cap) {                           // grow (rare): buffers are reused across uploads
            for (void *old : {(void *)ctx->well_i, (void *)ctx->well_j, (void *)ctx->well_q, (void *)ctx->well_md}) if (old) cudaFree(old);
            void *p;
            CUDA_TRY(ctx, cudaMalloc(&p, sizeof(int) * ctx->nwell)); ctx->well_i = (int *)p;
            CUDA_TRY(ctx, cudaMalloc(&p, sizeof(int) * ctx->nwell)); ctx->well_j = (int *)p;
            CUDA_TRY(ctx, cudaMalloc(&p, sizeof(double) * ctx->nwell)); ctx->well_q = (double *)p;
            CUDA_TRY(ctx, cudaMalloc(&p, sizeof(double) * ctx->nwell)); ctx->well_md = (double *)p;
            ctx->well_cap = ctx->nwell;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->well_i, ctx->hw_i, sizeof(int) * ctx->nwell, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->well_j, ctx->hw_j, sizeof(int) * ctx->nwell, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->well_q, ctx->hw_q, sizeof(double) * ctx->nwell, cudaMemcpyHostToDevice, ctx->stream));
#ifndef IMPEC_EMU
        if (!ctx->ev_wells) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_wells, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_wells, ctx->stream));
#endif
        ctx->bytes_h2d += (size_t)ctx->nwell * (2 * sizeof(int) + sizeof(double));
    }
    ctx->have_q = true;
    return IMPEC_OK;
}
extern "C" int impec_upload_q_rows(impec_ctx *ctx, const double *q) {
    const Grid &g = ctx->g;
    long long nw = 0;
    TRY(scan_wells(ctx, q, &nw));
    TRY(set_wells(ctx, q, nw));
    if ((long long)ctx->nwell * 16 > (long long)g.N1 * g.N2) {      // dense source term: plain copy
        const size_t w = (size_t)g.N2 * sizeof(double);
        CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->q + gidx(g, 0, 0), (size_t)g.ld * sizeof(double), q, w, w, (size_t)g.N1, cudaMemcpyHostToDevice, ctx->stream));
        ctx->bytes_h2d += w * (size_t)g.N1;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));          // the caller may reuse q at return (pageable memory is staged, pinned is not)
    } else {
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->q, 0, grid_elems(g) * sizeof(double), ctx->stream));
        if (ctx->nwell) LAUNCH(ctx, k_scatter_wells, dim3((unsigned)((ctx->nwell + 255) / 256)), dim3(256), ctx->nwell, ctx->well_i, ctx->well_j, ctx->well_q, g, ctx->q);
    }
    return IMPEC_OK;
}
extern "C" int impec_upload_q(impec_ctx *ctx, const double *q) { return impec_upload_q_rows(ctx, q + (size_t)ctx->g.i_off * ctx->g.N2); }
extern "C" int impec_set_q_dev_changed(impec_ctx *ctx) {
    const Grid &g = ctx->g;
    std::vector<double> q((size_t)g.N1 * g.N2);
    const size_t w = (size_t)g.N2 * sizeof(double);
    CUDA_TRY(ctx, cudaMemcpy2DAsync(q.data(), w, ctx->q + gidx(g, 0, 0), (size_t)g.ld * sizeof(double), w, (size_t)g.N1, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    long long nw = 0;
    TRY(scan_wells(ctx, q.data(), &nw));
    TRY(set_wells(ctx, q.data(), nw));
    return IMPEC_OK;
}
extern "C" int impec_upload_velocity(impec_ctx *ctx, const double *v_aos) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    const int H = g.N1 + 1 + 2 * ng, W = g.N2 + 1 + 2 * ng;
    const long long n = (long long)H * W;
    v_aos += (size_t)g.i_off * W * 2;                                          // this rank's rows of the global array
    ctx->dt_cached = false;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ta, v_aos, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->bytes_h2d += sizeof(double) * 2 * n;
    LAUNCH(ctx, k_deinterleave, dim3((unsigned)((n + 255) / 256)), dim3(256), g, ctx->ta, H, W, ctx->u, ctx->v, ng);
    return IMPEC_OK;
}
extern "C" int impec_download_velocity(impec_ctx *ctx, double *v_aos) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    const int H = g.N1 + 1 + 2 * ng, W = g.N2 + 1 + 2 * ng;
    const long long n = (long long)H * W;
    LAUNCH(ctx, k_interleave, dim3((unsigned)((n + 255) / 256)), dim3(256), g, ctx->ta, H, W, ctx->u, ctx->v, ng);
    v_aos += (size_t)g.i_off * W * 2;
    CUDA_TRY(ctx, cudaMemcpyAsync(v_aos, ctx->ta, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->bytes_d2h += sizeof(double) * 2 * n;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return IMPEC_OK;
}
// Whole-grid copy of a slab-decomposed field for logging: every rank contributes its rows, every rank receives all of them in the
// reference's host layout (ghosted c / p; AoS velocity).  Collective.  Ghost ROWS of the host array (and the wall face row u(N1,.)) are
// left as the caller has them; ghost columns come with the rows.  On one rank this is impec_download_*.
extern "C" int impec_gather_rows(impec_ctx *ctx, const char *field, double *host_full) {
    if (!field || !host_full) return fail(ctx, IMPEC_ERR_ARG, "impec_gather_rows: null argument");
    const std::string f(field);
    if (f != "c" && f != "p" && f != "v") return fail(ctx, IMPEC_ERR_ARG, "impec_gather_rows: field must be c, p or v");
    if (ctx->nranks == 1) return f == "c" ? impec_download_c(ctx, host_full) : (f == "p" ? impec_download_p(ctx, host_full) : impec_download_velocity(ctx, host_full));
    const Grid &g = ctx->g; const int ng = ctx->P.g, R = ctx->nranks, N1g = ctx->N1g;
    const size_t n = (size_t)g.N1 * g.ld, r0 = (size_t)IMPEC_GH * g.ld;
    const int nf = f == "v" ? 2 : 1;
    double *tmp = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void **)&tmp, sizeof(double) * n * R * nf));
    const double *src[2] = { f == "c" ? ctx->c : (f == "p" ? ctx->p : ctx->u), ctx->v };
    int rc = IMPEC_OK;
    for (int a = 0; a < nf && rc == IMPEC_OK; a++) rc = comm_allgather(ctx, src[a] + r0, tmp + (size_t)a * n * R, (long long)n);
    if (rc == IMPEC_OK) {
        if (nf == 1) {
            const size_t W = (size_t)g.N2 + 2 * ng, w = W * sizeof(double);
            if (cudaMemcpy2DAsync(host_full + (size_t)ng * W, w, tmp + (IMPEC_PJ - ng), (size_t)g.ld * sizeof(double), w, (size_t)N1g, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = IMPEC_ERR_CUDA;
            ctx->bytes_d2h += w * (size_t)N1g;
            if (rc == IMPEC_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = IMPEC_ERR_CUDA;
        } else {
            std::vector<double> hu(n * R), hv(n * R);
            if (cudaMemcpyAsync(hu.data(), tmp, sizeof(double) * n * R, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaMemcpyAsync(hv.data(), tmp + n * R, sizeof(double) * n * R, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = IMPEC_ERR_CUDA;
            ctx->bytes_d2h += 2 * sizeof(double) * n * R;
            const size_t W = (size_t)g.N2 + 1 + 2 * ng;
            if (rc == IMPEC_OK)
                for (int i = 0; i < N1g; i++)
                    for (int j = -ng; j < g.N2 + 1 + ng; j++) {
                        const size_t k = (size_t)i * g.ld + (size_t)(j + IMPEC_PJ), o = ((size_t)(i + ng) * W + (size_t)(j + ng)) * 2;
                        host_full[o] = hu[k]; host_full[o + 1] = hv[k];
                    }
        }
    }
    cudaFree(tmp);
    if (rc == IMPEC_ERR_CUDA) return fail(ctx, rc, "impec_gather_rows: copy failed");
    return rc;
}
extern "C" int impec_transfer_bytes(const impec_ctx *ctx, unsigned long long *h2d, unsigned long long *d2h) {
    if (h2d) *h2d = ctx->bytes_h2d; if (d2h) *d2h = ctx->bytes_d2h;
    return IMPEC_OK;
}
extern "C" int impec_field_layout(const impec_ctx *ctx, int *ld, int *row0, int *col0, int *rows) {
    if (ld) *ld = ctx->g.ld; if (row0) *row0 = IMPEC_GH; if (col0) *col0 = IMPEC_PJ; if (rows) *rows = ctx->g.rows;
    return IMPEC_OK;
}
extern "C" void *impec_field_dev(impec_ctx *ctx, const char *name) {
    const std::string k(name ? name : "");
    if (k == "c") { ctx->mob_valid = ctx->op_valid = false; return ctx->c; }
    if (k == "p") return ctx->p; if (k == "q") return ctx->q;
    if (k == "u") { ctx->dt_cached = false; return ctx->u; }
    if (k == "v") { ctx->dt_cached = false; return ctx->v; }
    // (diagnostic views for the tests: the pins' coarse-level shifts and the level arrays "AW0", "AV1", "DINV2", ...)
    if (k == "sigma") return ctx->sigma_dev;
    if (k.size() >= 3 && isdigit((unsigned char)k.back())) {
        const size_t l = (size_t)(k.back() - '0'); const std::string a = k.substr(0, k.size() - 1);
        if (l < ctx->L.size()) { Level &L = ctx->L[l]; if (a == "AW") return L.AW; if (a == "AS") return L.AS; if (a == "AV") return L.AV; if (a == "DINV") return L.DINV; }
    }
    return nullptr;
}

// ------------------------------------------------------------------------------------------------ a1 ghost fills
static int fill_scalar(impec_ctx *ctx, double *a, int periodic) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    const long long total = 2LL * ng * (g.N2 + 2 * ng) + 2LL * ng * g.N1;
    LAUNCH(ctx, k_fill_scalar, dim3((unsigned)((total + 255) / 256)), dim3(256), g, a, periodic, ng);
    return IMPEC_OK;
}
extern "C" int impec_fill_boundaries(impec_ctx *ctx) {
    if (ctx->P.is_test && ctx->nranks > 1) {
        // periodic domain on slabs (src/Util.cpp:215-256): the ghost columns wrap inside every row, the ghost rows are the rows of the
        // neighbouring rank with the ranks closed to a ring -- columns first, so that the rows travel with their corner cells
        Grid gp = ctx->g; gp.wall_lo = gp.wall_hi = 0;
        const int ng = ctx->P.g;
        const long long total = 2LL * ng * (gp.N2 + 2 * ng) + 2LL * ng * gp.N1;
        LAUNCH(ctx, k_fill_scalar, dim3((unsigned)((total + 255) / 256)), dim3(256), gp, ctx->c, 1, ng);
        return exchange_rows_periodic(ctx, ctx->c, ctx->g, ng);
    }
    TRY(fill_scalar(ctx, ctx->c, ctx->P.is_test ? 1 : 0));
    return exchange_rows(ctx, ctx->c, ctx->g, ctx->P.g);          // halo rows (with their ghost columns) from the neighbouring slabs
}

// ------------------------------------------------------------------------------------------------ a2-a8 operator
static LevelDev level_dev(const impec_ctx *ctx, int l, double mass) {
    const Level &L = ctx->L[l];
    LevelDev D;
    D.g = L.g; D.AW = L.AW; D.AS = L.AS; D.AV = L.AV; D.DINV = L.DINV; D.mass = mass;
    D.AWf = L.AWf; D.ASf = L.ASf; D.AVf = L.AVf; D.DINVf = L.DINVf;
    D.pin = L.pin; D.pin_end = L.pin_end; D.sig_idx = L.sig_idx; D.sig_end = L.sig_end; D.sig = ctx->sigma_dev + L.sig_off;
    D.cy0 = 0; D.cys = 1; D.pfd = 0;
    return D;
}

static int build_mobility(impec_ctx *ctx) {
    const Grid &g = ctx->g;
    LAUNCH(ctx, k_mobility, grid2d(g.N2 + 2, g.N1 + 2), kBlock, g, ctx->c, ctx->MU, ctx->D.mq, ctx->D.mu_res);
    if (ctx->P.nu != 1.) {
        if (ctx->P.is_test) LAUNCH(ctx, k_vertex<true>, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, ctx->MU, ctx->KV11, ctx->KV22, ctx->D);
        else LAUNCH(ctx, k_vertex<false>, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, ctx->MU, ctx->KV11, ctx->KV22, ctx->D);
    }
    ctx->mob_valid = true;
    return IMPEC_OK;
}

// FP64 DINV is read only by the per-colour (unfused) smoother and by the cycle kernels when they do not use the FP32 copies; slabs need
// it for the halo exchange.  Where nothing reads it, it is not written.
static bool need_dinv64(const impec_ctx *ctx) { return !ctx->links_f32 || !ctx->fused || ctx->nranks > 1; }
// DINV + FP32 copies of one level from its FP64 links (k_level_finish), halo rows of a slab by exchange
static int finish_level(impec_ctx *ctx, int l, const LevelDev &D, const double *sigma) {
    Level &L = ctx->L[l];
    const dim3 gr((unsigned)((L.g.ld + BX - 1) / BX), (unsigned)((L.g.rows + BY - 1) / BY));
    if (need_dinv64(ctx)) LAUNCH(ctx, k_level_finish<true>, gr, kBlock, D, sigma, L.DINV, L.AWf, L.ASf, L.AVf, L.DINVf);
    else LAUNCH(ctx, k_level_finish<false>, gr, kBlock, D, sigma, L.DINV, L.AWf, L.ASf, L.AVf, L.DINVf);
    if (L.dist) {
        TRY(exchange_rows(ctx, L.DINV, L.g, IMPEC_GH));
        for (int side = 0; side < 2; side++) {
            const int r0 = side == 0 ? 0 : L.g.N1 + IMPEC_GH, r1 = side == 0 ? IMPEC_GH : L.g.rows;
            const long long n = (long long)(r1 - r0) * L.g.ld;
            LAUNCH(ctx, k_f32_rows, dim3((unsigned)((n + 255) / 256)), dim3(256), L.g, r0, r1, (const double *)L.AW, (const double *)L.AS, (const double *)L.AV, (const double *)L.DINV,
                   L.AWf, L.ASf, L.AVf, L.DINVf);
        }
    }
    return IMPEC_OK;
}

extern "C" int impec_build_operator(impec_ctx *ctx) {
    const Grid &g = ctx->g;
    Level &L0 = ctx->L[0];
    const bool fused_build = ctx->build_fused && !ctx->P.is_test;
    if (fused_build) {
        // one march over c: links, inverse diagonal and FP32 copies of level 0 (MU / KV are not materialised: mob_valid stays false)
        BuildArgs A;
        A.g = g; A.c = ctx->c; A.AW = L0.AW; A.AS = L0.AS; A.AV = L0.AV; A.DINV = L0.DINV; A.AWf = L0.AWf; A.ASf = L0.ASf; A.AVf = L0.AVf; A.DINVf = L0.DINVf;
        A.pin = L0.pin; A.pin_end = L0.pin_end; A.P = ctx->D;
        const int warps = (g.N2 + BL_OWN - 1) / BL_OWN;
        const dim3 gb((unsigned)((warps + BL_BLOCK / 32 - 1) / (BL_BLOCK / 32)), (unsigned)((g.N1 + BL_ROWS - 1) / BL_ROWS));
        if (need_dinv64(ctx)) LAUNCH_COOP(ctx, k_build_links_march<true>, gb, dim3(BL_BLOCK), A, BL_ROWS);
        else LAUNCH_COOP(ctx, k_build_links_march<false>, gb, dim3(BL_BLOCK), A, BL_ROWS);
        ctx->mob_valid = false;
    } else {
        TRY(build_mobility(ctx));
        if (ctx->P.is_test) LAUNCH(ctx, k_links<true>, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, ctx->MU, ctx->KV11, ctx->KV22, L0.AW, L0.AS, L0.AV, ctx->D);
        else LAUNCH(ctx, k_links<false>, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, ctx->MU, ctx->KV11, ctx->KV22, L0.AW, L0.AS, L0.AV, ctx->D);
    }
    if (L0.pin_end - L0.pin > 1) {
        // A pinned ROW: the links between two Dirichlet cells carry nothing on the fine level (both ends are fixed) and must not be averaged
        // into the coarse operators either (the oracle's mg_build takes a fixed cell's links from its free neighbours' rows: 0 between two
        // fixed cells).  AS(1, j) couples (1, j-1) and (1, j).
        CUDA_TRY(ctx, cudaMemsetAsync(L0.AS + L0.pin, 0, (size_t)(L0.pin_end - L0.pin) * sizeof(double), ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(L0.ASf + L0.pin, 0, (size_t)(L0.pin_end - L0.pin) * sizeof(float), ctx->stream));
    }
    if (L0.dist) { TRY(exchange_rows(ctx, L0.AW, L0.g, IMPEC_GH)); TRY(exchange_rows(ctx, L0.AS, L0.g, IMPEC_GH)); TRY(exchange_rows(ctx, L0.AV, L0.g, IMPEC_GH)); }
    for (size_t l = 1; l < ctx->L.size(); l++) {
        Level &F = ctx->L[l - 1], &C = ctx->L[l];
        if ((int)l == ctx->Ld && ctx->nranks > 1) {
            // first replicated level: coarsen this rank's slab, then all-gather the rows into the global arrays
            LAUNCH(ctx, k_coarsen_links, grid2d(C.gl.N2 + 1, C.gl.N1 + 1), kBlock, F.g, F.AW, F.AS, F.AV, C.gl, C.AWl, C.ASl, C.AVl);
            const long long n = (long long)C.gl.N1 * C.gl.ld;
            const size_t r0 = (size_t)IMPEC_GH * C.gl.ld;
            TRY(comm_allgather(ctx, C.AWl + r0, C.AW + r0, n)); TRY(comm_allgather(ctx, C.ASl + r0, C.AS + r0, n)); TRY(comm_allgather(ctx, C.AVl + r0, C.AV + r0, n));
            continue;
        }
        LAUNCH(ctx, k_coarsen_links, grid2d(C.g.N2 + 1, C.g.N1 + 1), kBlock, F.g, F.AW, F.AS, F.AV, C.g, C.AW, C.AS, C.AV);
        if (C.dist) { TRY(exchange_rows(ctx, C.AW, C.g, IMPEC_GH)); TRY(exchange_rows(ctx, C.AS, C.g, IMPEC_GH)); TRY(exchange_rows(ctx, C.AV, C.g, IMPEC_GH)); }
    }
    if (ctx->rank == 0) {
        if (ctx->P.stype == IMPEC_HORIZONTAL) {
            // a pinned ROW: one shift per coarse column, level by level (same rule as the oracle's mg_build)
            for (size_t l = 1; l < ctx->L.size(); l++) {
                const int n = ctx->L[l].g.N2;
                if (l == 1) LAUNCH(ctx, k_pin_sigma_row, dim3((unsigned)((n + 255) / 256)), dim3(256), g, (const double *)L0.AW, (const double *)L0.AV, 1, n, ctx->sigma_dev + ctx->L[1].sig_off);
                else LAUNCH(ctx, k_sigma_coarsen, dim3((unsigned)((n + 255) / 256)), dim3(256), n, (const double *)(ctx->sigma_dev + ctx->L[l - 1].sig_off), ctx->sigma_dev + ctx->L[l].sig_off);
            }
        } else LAUNCH(ctx, k_pin_sigma, dim3(1), dim3(1), g, L0.AW, L0.AS, L0.AV, 1, 1, (int)ctx->L.size(), ctx->sigma_dev);
    }
    TRY(comm_bcast(ctx, ctx->sigma_dev, ctx->sigma_len, 0));
    for (size_t l = 0; l < ctx->L.size(); l++) {
        Level &L = ctx->L[l];
        if (l == 0 && fused_build) {
            if (L.dist) {           // the march wrote the interior; halo rows of DINV and of the FP32 copies
                TRY(exchange_rows(ctx, L.DINV, L.g, IMPEC_GH));
                for (int side = 0; side < 2; side++) {
                    const int r0 = side == 0 ? 0 : L.g.N1 + IMPEC_GH, r1 = side == 0 ? IMPEC_GH : L.g.rows;
                    const long long n = (long long)(r1 - r0) * L.g.ld;
                    LAUNCH(ctx, k_f32_rows, dim3((unsigned)((n + 255) / 256)), dim3(256), L.g, r0, r1, (const double *)L.AW, (const double *)L.AS, (const double *)L.AV, (const double *)L.DINV,
                           L.AWf, L.ASf, L.AVf, L.DINVf);
                }
            }
            continue;
        }
        TRY(finish_level(ctx, (int)l, level_dev(ctx, (int)l, 0.0), (const double *)(ctx->sigma_dev + L.sig_off)));
    }
    ctx->op_valid = true;
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ multigrid cycle
struct Cycle {
    impec_ctx *ctx;
    std::vector<LevelDev> D;
    int ncoarse, w_from, w_to;
    int nl;                       // levels used (1 = symmetric Gauss-Seidel preconditioner only)
};

static int smooth(Cycle &C, int l, const double *f, double *x, bool forward, bool zero) {
    impec_ctx *ctx = C.ctx; Level &L = ctx->L[l];
    const dim3 gr = grid2d(L.g.N2, L.g.N1);
    const int c0 = forward ? 0 : 1;
    if (zero) LAUNCH(ctx, k_smooth_color<true>, gr, kBlock, C.D[l], c0, f, x, L.t);
    else LAUNCH(ctx, k_smooth_color<false>, gr, kBlock, C.D[l], c0, f, x, L.t);
    LAUNCH(ctx, k_smooth_color<false>, gr, kBlock, C.D[l], 1 - c0, f, L.t, x);
    return IMPEC_OK;
}

// ---- fused path (kernels_fused.inl)
static const dim3 kFusedBlock(FT_W / 2, FT_NTY);
static const int kFusedSmem = FT_W * FT_H * (int)sizeof(double);
// k_mg_fused with the tile height chosen for the level (Level::th)
#define LAUNCH_FUSED(ctx, MODE, th, grid, ...) do { \
        const int smem_ = FT_W * (th) * (int)sizeof(double); \
        if ((ctx)->links_f32) { \
            if ((th) == 64) LAUNCH_COOP_SMEM(ctx, (k_mg_fused<MODE, true, 64>), grid, kFusedBlock, smem_, __VA_ARGS__); \
            else if ((th) == 32) LAUNCH_COOP_SMEM(ctx, (k_mg_fused<MODE, true, 32>), grid, kFusedBlock, smem_, __VA_ARGS__); \
            else LAUNCH_COOP_SMEM(ctx, (k_mg_fused<MODE, true, 16>), grid, kFusedBlock, smem_, __VA_ARGS__); \
        } else { \
            if ((th) == 64) LAUNCH_COOP_SMEM(ctx, (k_mg_fused<MODE, false, 64>), grid, kFusedBlock, smem_, __VA_ARGS__); \
            else if ((th) == 32) LAUNCH_COOP_SMEM(ctx, (k_mg_fused<MODE, false, 32>), grid, kFusedBlock, smem_, __VA_ARGS__); \
            else LAUNCH_COOP_SMEM(ctx, (k_mg_fused<MODE, false, 16>), grid, kFusedBlock, smem_, __VA_ARGS__); \
        } } while (0)
template <int MODE> static dim3 fused_grid(const Grid &g, int th = FT_H) {
    const int ti = th - 2 * (int)FusedGeom<MODE>::HI;
    return dim3((unsigned)((g.N2 + FT_TJ - 1) / FT_TJ), (unsigned)((g.N1 + ti - 1) / ti));
}
static bool use_march(const impec_ctx *ctx, const Level &L) { return ctx->march && L.march_rows > 0; }
static dim3 march_grid(const Level &L, int own) {
    const int warps = (L.g.N2 + own - 1) / own;
    return dim3((unsigned)((warps + MG_BLOCK / 32 - 1) / (MG_BLOCK / 32)), (unsigned)((L.g.N1 + L.march_rows - 1) / L.march_rows));
}
// PLAIN and POST read the iterate with a halo that belongs to neighbouring tiles, so they cannot run in place:
// they read L.x, write L.t and the two pointers are swapped (the level's current iterate is always L.x).
// xhalo > 0 (slab levels): exchange that many halo rows of the iterate first -- overlapped with the interior chunks on the marching path
static int fused_plain(Cycle &C, int l, const double *f, bool forward, int xhalo = 0) {
    impec_ctx *ctx = C.ctx; Level &L = ctx->L[l];
    if (use_march(ctx, L)) {
        ctx->march_launches++;
        const dim3 gm = march_grid(L, MGS_OWN);
        auto launch = [&](int y0, int ys, int ny) -> int {
            const dim3 gg(gm.x, (unsigned)ny); const LevelDev D = chunked(C.D[l], y0, ys, ctx->pf_post);
#define MPLAIN(F32, C0) LAUNCH_COOP(ctx, (k_mg_march_post<F32, false, C0, 3>), gg, dim3(MG_BLOCK), D, f, (const double *)L.x, L.t, L.g, (const double *)nullptr, (double *)nullptr, L.march_rows)
            if (ctx->links_f32) { if (forward) MPLAIN(true, 0); else MPLAIN(true, 1); } else { if (forward) MPLAIN(false, 0); else MPLAIN(false, 1); }
#undef MPLAIN
            return IMPEC_OK;
        };
        TRY(halo_overlapped(ctx, xhalo > 0 && L.dist, (int)gm.y, L.march_rows, L.g.N1, [&]() { return exchange_rows(ctx, L.x, L.g, xhalo); }, launch));
        std::swap(L.x, L.t);
        return IMPEC_OK;
    }
    if (xhalo > 0 && L.dist) TRY(exchange_rows(ctx, L.x, L.g, xhalo));
    LAUNCH_FUSED(ctx, FUSED_PLAIN, L.th, fused_grid<FUSED_PLAIN>(L.g, L.th), C.D[l], f, (const double *)L.x, L.t, forward ? 0 : 1, L.g,
                      (const double *)nullptr, (double *)nullptr, (double *)nullptr);
    std::swap(L.x, L.t);
    return IMPEC_OK;
}
// coarse-level view seen from a distributed level: a slab grid, and the matching window of a replicated array
static bool at_transition(const impec_ctx *ctx, int l) { return ctx->nranks > 1 && l + 1 == ctx->Ld; }
static const double *coarse_x_view(const impec_ctx *ctx, int l) {
    const Level &Lc = ctx->L[l + 1];
    return at_transition(ctx, l) ? Lc.x + (size_t)Lc.gl.i_off * Lc.gl.ld : Lc.x;
}
// On a slab level the IMPEC_HALO halo rows of f are exchanged first (overlapped with the interior chunks on the marching path)
static int fused_pre(Cycle &C, int l, const double *f) {
    impec_ctx *ctx = C.ctx; Level &L = ctx->L[l]; Level &Lc = ctx->L[l + 1];
    const bool tr = at_transition(ctx, l);
    if (use_march(ctx, L)) {
        ctx->march_launches++;
        const dim3 gm = march_grid(L, MGP_OWN);
        auto launch = [&](int y0, int ys, int ny) -> int {
            const dim3 gg(gm.x, (unsigned)ny); const LevelDev D = chunked(C.D[l], y0, ys, ctx->pf_pre);
#define MPRE(F32, PF, MB) LAUNCH_COOP(ctx, (k_mg_march_pre<F32, PF, MB>), gg, dim3(MG_BLOCK), D, f, L.x, tr ? Lc.gl : Lc.g, tr ? Lc.fl : Lc.f, L.march_rows)
            // rotating register slots (kernels_march.inl): 159 registers, 3 CTAs per SM, no spills.  (Capped at 128 registers for 4 CTAs per
            // SM the rotating form spills and loses what it gains: 1.75 ms against 1.58 at 16384^2, profiles/r2p_rotating_pre.txt.)
            if (!ctx->links_f32) MPRE(false, 4, 3); else MPRE(true, 2, 3);
#undef MPRE
            return IMPEC_OK;
        };
        return halo_overlapped(ctx, L.dist, (int)gm.y, L.march_rows, L.g.N1, [&]() { return exchange_rows(ctx, const_cast<double *>(f), L.g, IMPEC_HALO); }, launch);
    }
    if (L.dist) TRY(exchange_rows(ctx, const_cast<double *>(f), L.g, IMPEC_HALO));
    LAUNCH_FUSED(ctx, FUSED_PRE, L.th, fused_grid<FUSED_PRE>(L.g, L.th), C.D[l], f, (const double *)nullptr, L.x, 0, tr ? Lc.gl : Lc.g,
                      (const double *)nullptr, tr ? Lc.fl : Lc.f, (double *)nullptr);
    return IMPEC_OK;
}
// restricted right-hand side of the first replicated level: all-gather the slabs
static int gather_coarse_rhs(impec_ctx *ctx, int l) {
    if (!at_transition(ctx, l)) return IMPEC_OK;
    Level &Lc = ctx->L[l + 1];
    const size_t r0 = (size_t)IMPEC_GH * Lc.gl.ld;
    const long long n = (long long)Lc.gl.N1 * Lc.gl.ld;
#if !defined(IMPEC_EMU) && defined(IMPEC_WITH_NCCL)
    if (ctx->p2p_gather && !ctx->has_cb && (size_t)n * ctx->nranks * sizeof(double) <= ctx->p2p_win_bytes) return p2p_allgather(ctx, Lc.fl + r0, Lc.f + r0, n);
#endif
    return comm_allgather(ctx, Lc.fl + r0, Lc.f + r0, n);
}
// If the coarse level is a slab level too, 2 halo rows of its iterate are exchanged first (overlapped with the interior chunks on the
// marching path: only the first and the last chunk prolongate from coarse halo rows)
static int fused_post(Cycle &C, int l, const double *f, double *part, long long *nparts) {
    impec_ctx *ctx = C.ctx; Level &L = ctx->L[l]; Level &Lc = ctx->L[l + 1];
    if (use_march(ctx, L)) {
        ctx->march_launches++;
        const dim3 gm = march_grid(L, MGS_OWN);
        auto launch = [&](int y0, int ys, int ny) -> int {
            const dim3 gg(gm.x, (unsigned)ny); const LevelDev D = chunked(C.D[l], y0, ys, ctx->pf_post);
#define MPOST(F32, MB) LAUNCH_COOP(ctx, (k_mg_march_post<F32, true, 1, MB>), gg, dim3(MG_BLOCK), D, f, (const double *)L.x, L.t, at_transition(ctx, l) ? Lc.gl : Lc.g, coarse_x_view(ctx, l), part, L.march_rows)
            // 168 registers, 3 CTAs per SM.  (Forcing 4 CTAs per SM makes ptxas spill ~300 B per thread and DOUBLES the time.)
            if (!ctx->links_f32) MPOST(false, 3); else MPOST(true, 3);
#undef MPOST
            return IMPEC_OK;
        };
        TRY(halo_overlapped(ctx, Lc.dist, (int)gm.y, L.march_rows, L.g.N1, [&]() { return exchange_rows(ctx, Lc.x, Lc.g, 2); }, launch));
        std::swap(L.x, L.t);
        if (nparts) *nparts = (long long)gm.x * gm.y;
        return IMPEC_OK;
    }
    if (Lc.dist) TRY(exchange_rows(ctx, Lc.x, Lc.g, 2));
    const dim3 gr = fused_grid<FUSED_POST>(L.g, L.th);
    LAUNCH_FUSED(ctx, FUSED_POST, L.th, gr, C.D[l], f, (const double *)L.x, L.t, 1, at_transition(ctx, l) ? Lc.gl : Lc.g,
                      coarse_x_view(ctx, l), (double *)nullptr, part);
    std::swap(L.x, L.t);
    if (nparts) *nparts = (long long)gr.x * gr.y;
    return IMPEC_OK;
}

// One multigrid cycle for S x = f from a zero initial guess; the result is left in ctx->L[l].x (the fused path
// ping-pongs between the level's two buffers).  dot_part/dot_nparts (level 0, fused path): the last kernel also
// leaves the partial sums of <f,x> in dot_part.
static int mg_cycle_impl(Cycle &C, int l, const double *f, double *dot_part, long long *dot_nparts);
static int mg_cycle(Cycle &C, int l, const double *f, double *dot_part = nullptr, long long *dot_nparts = nullptr) {
    impec_ctx *ctx = C.ctx; (void)ctx;
#ifndef IMPEC_EMU
    // The replicated tail (levels >= tail_level) is the same launch sequence on the same buffers at every visit: it is
    // captured once into a CUDA graph and replayed, which removes the launch gaps of ~100 tiny kernels per visit.
    if (l == ctx->tail_level && ctx->use_graph && ctx->fused && C.nl == (int)ctx->L.size() && C.D[0].mass == 0.) {
        const int nl = (int)ctx->L.size();
        if (!ctx->graph_ready) {
            if (ctx->tail_exec) { cudaGraphExecDestroy(ctx->tail_exec); ctx->tail_exec = nullptr; }
            if (ctx->tail_graph) { cudaGraphDestroy(ctx->tail_graph); ctx->tail_graph = nullptr; }
            ctx->tail_init.clear(); ctx->tail_final.clear();
            for (int k = l; k < nl; k++) ctx->tail_init.push_back(std::make_pair(ctx->L[k].x, ctx->L[k].t));
            const long long l0 = ctx->launches;
            CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = mg_cycle_impl(C, l, f, nullptr, nullptr);
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &ctx->tail_graph);
            if (rc != IMPEC_OK) return rc;
            if (ce != cudaSuccess) return fail(ctx, IMPEC_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            CUDA_TRY(ctx, cudaGraphInstantiate(&ctx->tail_exec, ctx->tail_graph, 0));
            ctx->graph_launches = ctx->launches - l0; ctx->launches = l0;
            for (int k = l; k < nl; k++) ctx->tail_final.push_back(std::make_pair(ctx->L[k].x, ctx->L[k].t));
            ctx->graph_ready = true;
        }
        if (f != ctx->L[l].f) return fail(ctx, IMPEC_ERR_ARG, "multigrid tail called with a foreign right-hand side");
        CUDA_TRY(ctx, cudaGraphLaunch(ctx->tail_exec, ctx->stream));
        ctx->launches += ctx->graph_launches;
        for (int k = l; k < nl; k++) { ctx->L[k].x = ctx->tail_final[k - l].first; ctx->L[k].t = ctx->tail_final[k - l].second; }
        if (dot_nparts) *dot_nparts = 0;
        return IMPEC_OK;
    }
#endif
    return mg_cycle_impl(C, l, f, dot_part, dot_nparts);
}
static int mg_cycle_impl(Cycle &C, int l, const double *f, double *dot_part, long long *dot_nparts) {
    impec_ctx *ctx = C.ctx; Level &L = ctx->L[l];
    double *x = L.x;
    const int nl = C.nl;
    const bool fused = ctx->fused;
    if (dot_nparts) *dot_nparts = 0;
    if (!fused && ctx->nranks > 1) return fail(ctx, IMPEC_ERR_UNSUPPORTED, "the per-colour (mg_fused = 0) path is single-GPU only");
    if (fused && ctx->use_small && l > 0 && l < nl - 1 && l > C.w_to && !L.dist && nl - l <= SM_MAXLEV && (long long)L.g.N1 * L.g.N2 <= SM_MAXCELLS && C.D[l].mass == 0.) {
        // bottom of the hierarchy: the whole V-cycle over levels l .. nl-1 in one single-CTA launch
        SmallArgs A; A.nl = nl - l; A.ncoarse = C.ncoarse; A.xout = ctx->L[l].x;
        for (int k = l; k < nl; k++) A.D[k - l] = C.D[k];
        const size_t smem = small_smem_doubles(A) * sizeof(double);
        if (ctx->links_f32) LAUNCH_COOP_SMEM(ctx, k_mg_small<true>, dim3(1), dim3(SM_T), smem, A, f);
        else LAUNCH_COOP_SMEM(ctx, k_mg_small<false>, dim3(1), dim3(SM_T), smem, A, f);
        return IMPEC_OK;
    }
    if (l == nl - 1) {
        if (fused && !L.dist && L.g.N1 <= FT_H - 4 && L.g.N2 <= FT_TJ) {
            // the whole level fits one tile: all 2*ncoarse sweeps in a single launch (x0 = 0 inside the kernel)
            if (ctx->links_f32) LAUNCH_COOP_SMEM(ctx, k_mg_coarse_solve<true>, dim3(1, 1), kFusedBlock, kFusedSmem, C.D[l], f, L.x, 2 * C.ncoarse);
            else LAUNCH_COOP_SMEM(ctx, k_mg_coarse_solve<false>, dim3(1, 1), kFusedBlock, kFusedSmem, C.D[l], f, L.x, 2 * C.ncoarse);
            return IMPEC_OK;
        }
        if (fused && C.ncoarse == 1 && ctx->sgs_march && use_march(ctx, L)) {
            // one symmetric sweep from zero (the Crank-Nicolson preconditioner) on a big level: all four colour stages in ONE march.
            // On a slab the 3 halo rows of f make the sweep exact on the interior rows (no exchange of x between the two half sweeps).
            ctx->march_launches++;
            const dim3 gm = march_grid(L, MGP_OWN);
            auto launch = [&](int y0, int ys, int ny) -> int {
                const dim3 gg(gm.x, (unsigned)ny); const LevelDev D = chunked(C.D[l], y0, ys, ctx->pf_sgs);
                if (ctx->links_f32) LAUNCH_COOP(ctx, k_sgs_march<true>, gg, dim3(MG_BLOCK), D, f, L.x, dot_part, L.march_rows);
                else LAUNCH_COOP(ctx, k_sgs_march<false>, gg, dim3(MG_BLOCK), D, f, L.x, dot_part, L.march_rows);
                return IMPEC_OK;
            };
            TRY(halo_overlapped(ctx, L.dist, (int)gm.y, L.march_rows, L.g.N1, [&]() { return exchange_rows(ctx, const_cast<double *>(f), L.g, IMPEC_HALO); }, launch));
            if (dot_part && dot_nparts) *dot_nparts = (long long)gm.x * gm.y;
            return IMPEC_OK;
        }
        if (fused) {
            CUDA_TRY(ctx, cudaMemsetAsync(L.x, 0, grid_elems(L.g) * sizeof(double), ctx->stream));
            if (L.dist) TRY(exchange_rows(ctx, const_cast<double *>(f), L.g, IMPEC_HALO));
            for (int s = 0; s < C.ncoarse; s++) {
                TRY(fused_plain(C, l, f, true, s > 0 ? 2 : 0));          // (the first sweep starts from x = 0: nothing to exchange)
                TRY(fused_plain(C, l, f, false, 2));
            }
        } else
            for (int s = 0; s < C.ncoarse; s++) { TRY(smooth(C, l, f, x, true, s == 0)); TRY(smooth(C, l, f, x, false, false)); }
        return IMPEC_OK;
    }
    Level &Lc = ctx->L[l + 1];
    const int gamma = (l >= C.w_from && l <= C.w_to) ? 2 : 1;
    if (fused) {
        // On a slab level every kernel that reads a field with a halo is preceded by the exchange of exactly the rows it
        // reads (f: 3, x: 2); with the halos in place each rank computes the same values a single GPU would.
        // (the exchanges live in fused_pre / fused_post / fused_plain, which overlap them with the interior row chunks on the marching path)
        const bool dist = L.dist, tr = at_transition(ctx, l);
        TRY(fused_pre(C, l, f));                                      // [f: 3 halo rows] pre-smooth(0) + residual + restrict; writes x on 2 halo rows too
        TRY(gather_coarse_rhs(ctx, l));
        TRY(mg_cycle(C, l + 1, Lc.f));
        for (int gc = 1; gc < gamma; gc++) {
            TRY(fused_post(C, l, f, nullptr, nullptr));               // [coarse x: 2 halo rows] prolong + correct + backward sweep
            if (!dist && ctx->fused_mid) {
                // sweep + residual + restriction in one tile kernel (needs 4 halo rows: single-GPU / replicated levels only)
                LAUNCH_FUSED(ctx, FUSED_MID, L.th, fused_grid<FUSED_MID>(L.g, L.th), C.D[l], f, (const double *)L.x, L.t, 0, Lc.g, (const double *)nullptr, Lc.f, (double *)nullptr);
                std::swap(L.x, L.t);
            } else {
            TRY(fused_plain(C, l, f, true, 2));                       // [x: 2 halo rows]
            if (dist) TRY(exchange_rows(ctx, L.x, L.g, 2));
            LAUNCH(ctx, k_residual, grid2d(L.g.N2, L.g.N1 + 2), kBlock, C.D[l], f, (const double *)L.x, L.r);
            if (tr) LAUNCH(ctx, k_restrict, grid2d(Lc.gl.N2, Lc.gl.N1), kBlock, L.g, L.r, Lc.gl, Lc.fl);
            else LAUNCH(ctx, k_restrict, grid2d(Lc.g.N2, Lc.g.N1), kBlock, L.g, L.r, Lc.g, Lc.f);
            }
            TRY(gather_coarse_rhs(ctx, l));
            TRY(mg_cycle(C, l + 1, Lc.f));
        }
        TRY(fused_post(C, l, f, dot_part, dot_nparts));               // prolong + correct + post-smooth (+ <f,x>)
        return IMPEC_OK;
    }
    TRY(smooth(C, l, f, x, true, true));
    for (int gc = 0; gc < gamma; gc++) {
        LAUNCH(ctx, k_residual, grid2d(L.g.N2, L.g.N1), kBlock, C.D[l], f, x, L.r);
        LAUNCH(ctx, k_restrict, grid2d(Lc.g.N2, Lc.g.N1), kBlock, L.g, L.r, Lc.g, Lc.f);
        TRY(mg_cycle(C, l + 1, Lc.f));
        LAUNCH(ctx, k_prolong_add, grid2d(L.g.N2, L.g.N1), kBlock, Lc.g, Lc.x, L.g, x, L.pin, L.pin_end);
        if (gc < gamma - 1) { TRY(smooth(C, l, f, x, false, false)); TRY(smooth(C, l, f, x, true, false)); }
    }
    TRY(smooth(C, l, f, x, false, false));
    return IMPEC_OK;
}

static int make_cycle(impec_ctx *ctx, Cycle &C, double mass, int nlevels_used) {
    C.ctx = ctx;
    C.D.clear();
    for (size_t l = 0; l < ctx->L.size(); l++) { LevelDev D = level_dev(ctx, (int)l, mass); if (mass != 0.) { D.pin = D.pin_end = -1; D.sig_idx = D.sig_end = -1; } C.D.push_back(D); }
    C.ncoarse = ctx->P.mg_ncoarse > 0 ? ctx->P.mg_ncoarse : 8;
    C.w_from = ctx->P.mg_w_from; C.w_to = ctx->P.mg_w_to;
    C.nl = nlevels_used > 0 ? nlevels_used : (int)ctx->L.size();
    return IMPEC_OK;
}

extern "C" int impec_vcycle(impec_ctx *ctx, const double *r_host, double *z_host) {
    if (!ctx->op_valid) TRY(impec_build_operator(ctx));
    Cycle C; TRY(make_cycle(ctx, C, 0.0, 0));
    const Grid &g = ctx->g;
    const size_t w = (size_t)g.N2 * sizeof(double);
    if (r_host) {
        CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->r + gidx(g, 0, 0), (size_t)g.ld * sizeof(double), r_host + (size_t)g.i_off * g.N2, w, w, (size_t)g.N1, cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->L[0].pin >= 0) LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->r + ctx->L[0].pin, 0.0);
    }
    TRY(mg_cycle(C, 0, ctx->r));
    if (z_host) {
        CUDA_TRY(ctx, cudaMemcpy2DAsync(z_host + (size_t)g.i_off * g.N2, w, ctx->L[0].x + gidx(g, 0, 0), (size_t)g.ld * sizeof(double), w, (size_t)g.N1, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ a8-a11 pressure solve
// src/EllipticSolver.cpp:357-410 / src/MGSolver.cpp:524-545: in test mode the pin's coefficient is chosen so
// that p(1,1) equals the manufactured solution
static double q_test_host(const impec_params &P, double x, double y) {
    const double t = 2.;
    switch (P.testno) {
    case 1: return 2. / P.mu_res * cos(x) * cos(y);
    case 2: { const double c1 = -0.3, c2 = -0.3, a1 = exp(c1 * x), a2 = exp(c2 * y), a1_x = c1 * exp(c1 * x), a2_y = c2 * exp(c2 * y);
              return 1. / P.mu_res * ((a1 + a2) * cos(x) * cos(y) + a1_x * sin(x) * cos(y) + a2_y * cos(x) * sin(y)); }
    case 3: { const double a1 = 2. + sin(t * x) * sin(y), a2 = 2. + sin(t * x) * sin(y), a1_x = t * cos(t * x) * sin(y), a2_y = cos(y) * sin(t * x);
              return 1. / P.mu_res * ((a1 + a2) * cos(x) * cos(y) + a1_x * sin(x) * cos(y) + a2_y * cos(x) * sin(y)); }
    case 4: { const double a1 = 2. + sin(t * x) * sin(y), a2 = 2. + sin(x) * sin(t * y), a1_x = t * cos(t * x) * sin(y), a2_y = t * cos(t * y) * sin(x);
              return 1. / P.mu_res * ((a1 + a2) * cos(x) * cos(y) + a1_x * sin(x) * cos(y) + a2_y * cos(x) * sin(y)); }
    default: return 0.;
    }
}

// PCG exactly as HYPRE's struct PCG is configured upstream (src/MGSolver.cpp:597-603: TwoNorm, RelChange off, zero
// initial guess): stop when <r,r>/<b,b> < eps^2 with r the RECURRENCE residual.  On entry: rhs and x are set, the
// partial sums of <b,b> are already reduced into scal[S_BIPROD].  The preconditioner is one multigrid cycle C.
// x0_pin_only: the caller guarantees that x is zero except at the pinned cell (the reference's x0 = 0): the initial residual is
// then a copy of rhs plus a 3 x 3 fix-up around the pin instead of a full residual pass (bit-identical)
static int pcg_solve(impec_ctx *ctx, Cycle &C, const double *rhs, double *x, double eps, impec_solve_info *info, bool x0_pin_only = false) {
    const Grid &g = ctx->g;
    const dim3 gr = grid2d(g.N2, g.N1);
    const long long nparts = (long long)gr.x * gr.y;
    const LevelDev &A = C.D[0];
    const dim3 gres = grid2d(g.N2, g.N1 + 2);                    // k_residual also fills one halo row per neighbouring slab
    // (a pinned row -- source_type horizontal -- couples into a whole row of neighbours: full residual pass; the choice must be the same on
    //  every rank, so it keys on the configuration, not on this rank's share of the pins)
    if (x0_pin_only && ctx->fast_r0 && ctx->P.stype != IMPEC_HORIZONTAL) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->r, rhs, grid_elems(g) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        if (A.pin >= 0) LAUNCH(ctx, k_residual_pin_fix, dim3(1), dim3(32), A, rhs, (const double *)x, ctx->r);
    } else {
        TRY(exchange_rows(ctx, x, g, 2));
        LAUNCH(ctx, k_residual, gres, kBlock, A, rhs, (const double *)x, ctx->r);
    }
    LAUNCH_COOP(ctx, k_dot, gr, kBlock, g, ctx->r, ctx->r, ctx->part);
    TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_IPROD));
    TRY(fetch_scalars(ctx, S_IPROD, 3));                         // S_IPROD .. S_BIPROD
    const double bi_prod = ctx->h_scal[S_BIPROD];
    double i_prod = ctx->h_scal[S_IPROD];
    int it = 0, converged = 0;
    double last_true = -1.;
    if (!(bi_prod > 0.)) converged = 1;                          // b = 0  =>  x = 0 (HYPRE returns immediately)
    else if (i_prod / bi_prod < eps * eps) converged = 1;
    int ig = S_GAMMA0;                                           // which scalar slot holds the current gamma
    long long dparts = 0;
    // Fused path: d <- z + beta d, s <- S d and <s,d> in ONE tile kernel (k_pcg_matvec); the search direction ping-pongs
    // between two buffers because the kernel reads d with a halo and writes its interior.
    const bool fuse_dir = ctx->fused;
    const dim3 gux = update_xr_grid(g);
    const int mvrows = matvec_march_rows(g);
    const dim3 gmm = matvec_march_grid(g);
    const dim3 gmv = fused_grid<FUSED_PLAIN>(g);
    const long long nmv = (long long)gmv.x * gmv.y;
    double *dcur = ctx->d, *dalt = ctx->d2buf;
    bool first_dir = true;
    // deferred update of the solution: needs the two direction buffers of the fused path (d_k stays intact while d_{k+1} is built)
    const bool defer_x = ctx->defer_x && fuse_dir;
    bool x_pending = false;
    auto flush_x = [&]() -> int {
        if (x_pending) { LAUNCH(ctx, k_flush_x, gux, dim3(UX_BLOCK), g, (const double *)(ctx->scal + S_ALPHA_PREV), (const double *)dcur, x); x_pending = false; }
        return IMPEC_OK;
    };
    if (!converged) {
        TRY(mg_cycle(C, 0, ctx->r, ctx->part, &dparts));
        if (!fuse_dir) LAUNCH(ctx, k_update_d<true>, gr, kBlock, g, ctx->scal, (const double *)ctx->L[0].x, dcur);
        if (!dparts) { LAUNCH_COOP(ctx, k_dot, gr, kBlock, g, ctx->r, ctx->L[0].x, ctx->part); dparts = nparts; }
        TRY(reduce_final(ctx, RED_SUM, dparts, 1, 0, ig));
    }
    while (!converged && it < ctx->P.max_iter) {
        it++;
        if (fuse_dir) {
            // beta = gamma_new / gamma_old sits in scal[S_BETA_NUM], scal[S_BETA_DEN] (set at the end of the previous pass)
            if (ctx->matvec_march) {
                // [z and d: 1 halo row each, one exchange] overlapped with the interior row chunks
                auto launch = [&](int y0, int ys, int ny) -> int {
                    const dim3 gg(gmm.x, (unsigned)ny); const LevelDev D = chunked(A, y0, ys, ctx->pf_mv);
                    if (first_dir) LAUNCH_COOP(ctx, k_pcg_matvec_march<true>, gg, dim3(MV_BLOCK), D, (const double *)(ctx->scal + S_BETA_NUM), (const double *)ctx->L[0].x, (const double *)dcur, dalt, ctx->s, ctx->part, mvrows);
                    else LAUNCH_COOP(ctx, k_pcg_matvec_march<false>, gg, dim3(MV_BLOCK), D, (const double *)(ctx->scal + S_BETA_NUM), (const double *)ctx->L[0].x, (const double *)dcur, dalt, ctx->s, ctx->part, mvrows);
                    return IMPEC_OK;
                };
                TRY(halo_overlapped(ctx, ctx->L[0].dist, (int)gmm.y, mvrows, g.N1, [&]() { return exchange_rows2(ctx, ctx->L[0].x, first_dir ? nullptr : dcur, g, 1); }, launch));
            } else {
                TRY(exchange_rows2(ctx, ctx->L[0].x, first_dir ? nullptr : dcur, g, 1));
                if (first_dir) LAUNCH_COOP_SMEM(ctx, k_pcg_matvec<true>, gmv, kFusedBlock, kFusedSmem, A, (const double *)(ctx->scal + S_BETA_NUM), (const double *)ctx->L[0].x, (const double *)dcur, dalt, ctx->s, ctx->part);
                else LAUNCH_COOP_SMEM(ctx, k_pcg_matvec<false>, gmv, kFusedBlock, kFusedSmem, A, (const double *)(ctx->scal + S_BETA_NUM), (const double *)ctx->L[0].x, (const double *)dcur, dalt, ctx->s, ctx->part);
            }
            std::swap(dcur, dalt); first_dir = false;
            TRY(reduce_final(ctx, RED_SUM, ctx->matvec_march ? (long long)gmm.x * gmm.y : nmv, 1, 0, S_SDOTP));
        } else {
            TRY(exchange_rows(ctx, dcur, g, 1));
            LAUNCH_COOP(ctx, k_matvec_dot, gr, kBlock, A, dcur, ctx->s, ctx->part);
            TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_SDOTP));
        }
        // alpha = gamma / <s,d>: the kernel reads scal[0]/scal[1], so gamma must sit in slot 0
        if (ig != S_GAMMA0) { CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal + S_GAMMA0, ctx->scal + ig, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)); ig = S_GAMMA0; }
        if (defer_x) {
            // deferred solution update (k_update_xr MODE 1 / 2): odd iterations touch r only, even ones add both pending directions to x
            if (!x_pending) LAUNCH_COOP(ctx, k_update_xr<1>, gux, dim3(UX_BLOCK), g, ctx->scal, (const double *)dcur, ctx->s, x, ctx->r, ctx->part, (const double *)nullptr, ctx->scal + S_ALPHA_PREV);
            else LAUNCH_COOP(ctx, k_update_xr<2>, gux, dim3(UX_BLOCK), g, ctx->scal, (const double *)dcur, ctx->s, x, ctx->r, ctx->part, (const double *)dalt, ctx->scal + S_ALPHA_PREV);
            x_pending = !x_pending;
        } else
            LAUNCH_COOP(ctx, k_update_xr<0>, gux, dim3(UX_BLOCK), g, ctx->scal, (const double *)dcur, ctx->s, x, ctx->r, ctx->part, (const double *)nullptr, (double *)nullptr);
        TRY(reduce_final(ctx, RED_SUM, (long long)gux.x * gux.y, 1, 0, S_IPROD));
        TRY(fetch_scalars(ctx, S_IPROD, 1));
        i_prod = ctx->h_scal[S_IPROD];
        if (i_prod / bi_prod < eps * eps) { converged = 1; break; }
        if (!(i_prod == i_prod)) return fail(ctx, IMPEC_ERR_NOCONV, "PCG produced NaN at iteration %d", it);
        if (ctx->stagnation_guard && it >= 8 && it % 3 == 2) {
            // Optional: eps = 1e-14 is below what FP64 can deliver on large grids; the recurrence residual keeps
            // falling while b - A x does not.  Every third iteration compare the two; stop when the true residual is an
            // order of magnitude above the recurrence value and no longer improving.
            TRY(flush_x());
            TRY(exchange_rows(ctx, x, g, 2));
            LAUNCH(ctx, k_residual, gres, kBlock, A, rhs, (const double *)x, ctx->s);
            LAUNCH_COOP(ctx, k_dot, gr, kBlock, g, ctx->s, ctx->s, ctx->part);
            TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_TMP));
            TRY(fetch_scalars(ctx, S_TMP, 1));
            const double true_prod = ctx->h_scal[S_TMP];
            if (true_prod > 100. * i_prod && last_true > 0. && true_prod > 0.25 * last_true) { converged = 2; break; }
            last_true = true_prod;
        }
        TRY(mg_cycle(C, 0, ctx->r, ctx->part, &dparts));
        if (!dparts) { LAUNCH_COOP(ctx, k_dot, gr, kBlock, g, ctx->r, ctx->L[0].x, ctx->part); dparts = nparts; }
        TRY(reduce_final(ctx, RED_SUM, dparts, 1, 0, S_BETA_NUM));
        // beta = gamma_new / gamma_old ; d = z + beta d      (k_update_d reads scal'[0]/scal'[1])
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scal + S_BETA_DEN, ctx->scal + S_GAMMA0, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        if (!fuse_dir) LAUNCH(ctx, k_update_d<false>, gr, kBlock, g, ctx->scal + S_BETA_NUM, (const double *)ctx->L[0].x, dcur);
        ig = S_BETA_NUM;
    }
    // true residual with the final x
    TRY(flush_x());
    TRY(exchange_rows(ctx, x, g, 2));
    { const int rrows = 64;
      const dim3 grn((unsigned)((g.N2 + BX - 1) / BX), (unsigned)((g.N1 + rrows - 1) / rrows));
      LAUNCH_COOP(ctx, k_residual_norm, grn, kBlock, A, rhs, (const double *)x, ctx->part, rrows);
      TRY(reduce_final(ctx, RED_SUM, (long long)grn.x * grn.y, 1, 0, S_TMP)); }
    TRY(fetch_scalars(ctx, S_TMP, 1));
    if (info) {
        info->iterations = it; info->converged = converged;
        info->relres = bi_prod > 0. ? sqrt(i_prod / bi_prod) : 0.;
        info->true_relres = bi_prod > 0. ? sqrt(ctx->h_scal[S_TMP] / bi_prod) : 0.;
    }
    if (info && converged == 2) info->converged = 2;          // stopped by the stagnation guard at the attainable accuracy
    // Not converged within max_iter: the reference never checks HYPRE's convergence and carries on with the last iterate
    // (src/MGSolver.cpp:612-622), so this is a WARNING (positive code), reported through info->converged = 0 as well.
    return converged ? IMPEC_OK : fail(ctx, IMPEC_WARN_NOCONV, "PCG did not reach eps=%g in %d iterations (relres %g)", eps, it, sqrt(i_prod / bi_prod));
}

extern "C" int impec_pressure_solve(impec_ctx *ctx, double eps, impec_solve_info *info) {
    if (!ctx->have_q) return fail(ctx, IMPEC_ERR_ARG, "pressure solve before impec_upload_q");
    const Grid &g = ctx->g;
    const dim3 gr = grid2d(g.N2, g.N1);
    const long long nparts = (long long)gr.x * gr.y;
    TRY(impec_build_operator(ctx));
    Cycle C; TRY(make_cycle(ctx, C, 0.0, 0));
    const LevelDev &A = C.D[0];
    // pinned value p(1,1) = b/(kappa/dA) = -q(1,1) dA / kappa       (src/MGSolver.cpp:517-567)
    const double dA = ctx->D.dx * ctx->D.dy;
    double kappa = 1.;
    if (ctx->P.is_test) {
        const double x = ctx->P.x_lb + 1.5 * ctx->D.dx, y = ctx->P.y_lb + 1.5 * ctx->D.dy;
        kappa = -dA * q_test_host(ctx->P, x, y) / (cos(x) * cos(y));
    }
    const double p_pin = -ctx->q_pin / (kappa / dA);
    const double pin_scale = ctx->P.stype == IMPEC_HORIZONTAL ? -dA : 0.;        // a pinned row: p(1,j) = -q(1,j) dA  (kappa = 1)
    double *x = ctx->p;
    if (ctx->warm_start) {
        // keep the previous pressure as initial guess (not the reference's behaviour: x0 = 0, MGSolver.cpp:577)
        LAUNCH(ctx, k_copy_interior, gr, kBlock, g, ctx->p, ctx->d);
        LAUNCH_COOP(ctx, k_setup_rhs, gr, kBlock, g, ctx->q, ctx->rhs, x, A.pin, A.pin_end, p_pin, pin_scale, ctx->part);
        LAUNCH(ctx, k_copy_interior, gr, kBlock, g, ctx->d, x);
        if (A.pin >= 0) LAUNCH(ctx, k_set_pins, dim3((unsigned)((A.pin_end - A.pin + 255) / 256)), dim3(256), x, (const double *)ctx->q, A.pin, A.pin_end, p_pin, pin_scale);      // only the rank that owns the pinned cells
    } else {
        LAUNCH_COOP(ctx, k_setup_rhs, gr, kBlock, g, ctx->q, ctx->rhs, x, A.pin, A.pin_end, p_pin, pin_scale, ctx->part);
    }
    TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_BIPROD));
    return pcg_solve(ctx, C, ctx->rhs, x, eps, info, !ctx->warm_start);
}

// ------------------------------------------------------------------------------------------------ a12 velocities
extern "C" int impec_calc_velocities(impec_ctx *ctx) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    ctx->dt_cached = false;
    TRY(fill_scalar(ctx, ctx->p, 0));                               // src/EllipticSolver.cpp:663
    TRY(exchange_rows(ctx, ctx->p, g, ng));
    if (ctx->vel_march && !ctx->P.is_test) {
        // one march from c and p: u, v and the reductions impec_max_stable_dt needs (kernels_build.inl)
        VelArgs A;
        A.g = g; A.c = ctx->c; A.p = ctx->p; A.u = ctx->u; A.v = ctx->v; A.part = ctx->part; A.stride = ctx->part_cap; A.P = ctx->D;
        const int warps = (g.N2 + 1 + BL_OWN - 1) / BL_OWN;
        const dim3 gb((unsigned)((warps + BL_BLOCK / 32 - 1) / (BL_BLOCK / 32)), (unsigned)((g.N1 + BL_ROWS - 1) / BL_ROWS));
        const long long nparts = (long long)gb.x * gb.y;
        if (ctx->P.diffusion) LAUNCH_COOP(ctx, k_velocity_march<true>, gb, dim3(BL_BLOCK), A, BL_ROWS);
        else LAUNCH_COOP(ctx, k_velocity_march<false>, gb, dim3(BL_BLOCK), A, BL_ROWS);
        TRY(reduce_final(ctx, RED_MAX, nparts, 1, 0, S_DT));
        if (ctx->P.diffusion) {
            LAUNCH_COOP(ctx, k_reduce_final<RED_MIN>, dim3(1), dim3(1024), ctx->part + ctx->part_cap, nparts, 0LL, ctx->scal + S_DT + 1);
            TRY(comm_allreduce(ctx, ctx->scal + S_DT + 1, 1, RED_MIN));
        }
        ctx->dt_cached = true;
    } else {
        if (!ctx->mob_valid) TRY(build_mobility(ctx));
        if (ctx->P.is_test) LAUNCH(ctx, k_velocity<true>, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, ctx->p, ctx->MU, ctx->KV11, ctx->KV22, ctx->u, ctx->v, ctx->D);
        else LAUNCH(ctx, k_velocity<false>, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, ctx->p, ctx->MU, ctx->KV11, ctx->KV22, ctx->u, ctx->v, ctx->D);
    }
    const long long n = (long long)(g.N1 + 1 + 2 * ng) * (g.N2 + 1 + 2 * ng);
    LAUNCH(ctx, k_fill_velocity, dim3((unsigned)((n + 255) / 256)), dim3(256), g, ctx->u, ctx->v, ng);   // :682
    TRY(exchange_rows(ctx, ctx->u, g, ng + 1));                     // u row N1 is the neighbour's row 0; N1+1, N1+2 are its rows 1, 2
    TRY(exchange_rows(ctx, ctx->v, g, ng));
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ a13 time step
extern "C" int impec_max_stable_dt(impec_ctx *ctx, double *dt) {
    const Grid &g = ctx->g;
    const dim3 gr = grid2d(g.N2, g.N1);
    const long long nparts = (long long)gr.x * gr.y;
    const impec_params &P = ctx->P;
    if (P.is_test) {                                                // src/HyperbolicSolver.cpp:100-121
        LAUNCH_COOP(ctx, k_dt_reduce<1>, gr, kBlock, g, ctx->u, ctx->v, ctx->D, ctx->part);
        TRY(reduce_final(ctx, RED_MIN, nparts, 1, 0, S_DT));
        TRY(fetch_scalars(ctx, S_DT, 1));
        *dt = ctx->h_scal[S_DT] * P.cfl;
        return IMPEC_OK;
    }
    if (!ctx->dt_cached) {                                          // (calc_velocities reduces on the fly: nothing to launch then)
        LAUNCH_COOP(ctx, k_dt_reduce<0>, gr, kBlock, g, ctx->u, ctx->v, ctx->D, ctx->part);
        TRY(reduce_final(ctx, RED_MAX, nparts, 1, 0, S_DT));
        if (P.diffusion) {
            LAUNCH_COOP(ctx, k_dt_reduce<2>, gr, kBlock, g, ctx->u, ctx->v, ctx->D, ctx->part);
            TRY(reduce_final(ctx, RED_MIN, nparts, 1, 0, S_DT + 1));
        }
    }
    TRY(fetch_scalars(ctx, S_DT, 2));
    const double max_vel = ctx->h_scal[S_DT];
    const double a = ctx->D.dx / max_vel, b = ctx->D.dy / max_vel;
    double dt_adv = (b < a ? b : a);
    dt_adv *= P.cfl;
    double dt_diff = DBL_MAX;
    if (P.diffusion) { const double mn = ctx->h_scal[S_DT + 1] < 1.e10 ? ctx->h_scal[S_DT + 1] : 1.e10; dt_diff = mn * P.dt_diff_limit; }
    *dt = (P.ptype == IMPEC_CRANKNICOLSON) ? dt_adv : (dt_diff < dt_adv ? dt_diff : dt_adv);
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ a15 source
extern "C" int impec_source_split(impec_ctx *ctx, double dt) {
    if (!ctx->have_q) return fail(ctx, IMPEC_ERR_ARG, "source step before impec_upload_q");
    if (ctx->nwell == 0) {
        if (ctx->nranks == 1) return IMPEC_OK;
        LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->scal + S_SRC, 0.0);      // a slab without wells still joins the sum
        TRY(comm_allreduce(ctx, ctx->scal + S_SRC, 1, RED_SUM));
        LAUNCH(ctx, k_accum_abs, dim3(1), dim3(1), ctx->scal + S_SRC, ctx->scal + S_OUTFLUX);
        return IMPEC_OK;
    }
    ctx->mob_valid = ctx->op_valid = false;
    LAUNCH(ctx, k_source_tr, dim3((unsigned)((ctx->nwell + 255) / 256)), dim3(256), ctx->nwell, ctx->well_i, ctx->well_j, ctx->well_q, ctx->g, ctx->c,
           dt, ctx->P.inj_fluid_conc, ctx->well_md);
    LAUNCH_COOP(ctx, k_reduce_final<RED_SUM>, dim3(1), dim3(1024), ctx->well_md, (long long)ctx->nwell, 0LL, ctx->scal + S_SRC);
    TRY(comm_allreduce(ctx, ctx->scal + S_SRC, 1, RED_SUM));
    LAUNCH(ctx, k_accum_abs, dim3(1), dim3(1), ctx->scal + S_SRC, ctx->scal + S_OUTFLUX);
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ a16 advection
template <int DIR>
static int sweep(impec_ctx *ctx, const double *cin, double *cout, const double *other, double dt, int i0, int i1, int j0, int j1) {
    const Grid &g = ctx->g;
    const double *vel = DIR == 0 ? ctx->u : ctx->v;
    const double h = DIR == 0 ? ctx->D.dx : ctx->D.dy;
    const double phi = ctx->P.porosity;
    const bool phi1 = (phi == 1.0), up = ctx->P.hyper_method == IMPEC_UPWIND, avg = other != nullptr;
    if (!up && ctx->sweep_shared) {
        // MUSCL with every reconstruction evaluated once (k_muscl_x marches down the rows, k_muscl_y shares through smem)
        if (DIR == 0) {
            const int jstart = j0 - ((j0 + IMPEC_PJ) & 3);            // 32-byte aligned first column
            const dim3 gx((unsigned)((j1 - jstart + SX_BLOCK - 1) / SX_BLOCK), (unsigned)((i1 - i0 + SX_ROWS - 1) / SX_ROWS));
#define SWX(PH, AV) do { auto kern = k_muscl_x<PH, AV>; LAUNCH(ctx, kern, gx, dim3(SX_BLOCK), g, cin, vel, cout, other, dt, h, phi, i0, i1, j0, j1, jstart, ctx->pf_sweep); } while (0)
            if (phi1) { if (avg) { SWX(true, true); } else { SWX(true, false); } } else { if (avg) { SWX(false, true); } else { SWX(false, false); } }
#undef SWX
        } else if (ctx->sweep_y_pairs && ((j0 + IMPEC_PJ) & 1) == 0) {
            // column pairs per lane, neighbours by shuffle, no shared memory (k_muscl_y2)
            const int warps = (j1 - j0 + SY2_OWN - 1) / SY2_OWN;
            const dim3 gy((unsigned)((warps + SY2_BLOCK / 32 - 1) / (SY2_BLOCK / 32)), (unsigned)((i1 - i0 + SY2_ROWS - 1) / SY2_ROWS));
#define SWY2(PH, AV) do { auto kern = k_muscl_y2<PH, AV>; LAUNCH_COOP(ctx, kern, gy, dim3(SY2_BLOCK), g, cin, vel, cout, other, dt, h, phi, i0, i1, j0, j1, ctx->pf_sweep); } while (0)
            if (phi1) { if (avg) { SWY2(true, true); } else { SWY2(true, false); } } else { if (avg) { SWY2(false, true); } else { SWY2(false, false); } }
#undef SWY2
        } else {
            const dim3 gy((unsigned)((j1 - j0 + SY_BLOCK - 1) / SY_BLOCK), (unsigned)((i1 - i0 + SY_ROWS - 1) / SY_ROWS));
#define SWY(PH, AV) do { auto kern = k_muscl_y<PH, AV>; LAUNCH_COOP(ctx, kern, gy, dim3(SY_BLOCK), g, cin, vel, cout, other, dt, h, phi, i0, i1, j0, j1); } while (0)
            if (phi1) { if (avg) { SWY(true, true); } else { SWY(true, false); } } else { if (avg) { SWY(false, true); } else { SWY(false, false); } }
#undef SWY
        }
        return IMPEC_OK;
    }
    const dim3 gr = grid2d(j1 - j0, i1 - i0);
#define SW(M, PH, AV) do { auto kern = k_sweep<DIR, M, PH, AV>; LAUNCH(ctx, kern, gr, kBlock, g, cin, vel, cout, other, dt, h, phi, i0, i1, j0, j1); } while (0)
    if (!up) { if (phi1) { if (avg) { SW(0, true, true); } else { SW(0, true, false); } } else { if (avg) { SW(0, false, true); } else { SW(0, false, false); } } }
    else     { if (phi1) { if (avg) { SW(1, true, true); } else { SW(1, true, false); } } else { if (avg) { SW(1, false, true); } else { SW(1, false, false); } } }
#undef SW
    return IMPEC_OK;
}

static int copy_field(impec_ctx *ctx, const double *src, double *dst) {
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, grid_elems(ctx->g) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return IMPEC_OK;
}

// Reference semantics of a single sweep (src/HyperbolicSplitSolver.cpp:113-159): the x sweep updates
// i in [0,N1) for ALL j in [-g,N2+g); the y sweep updates ALL i in [-g,N1+g) for j in [0,N2); other cells keep
// their value.  In place on c (through the temporary ta).
extern "C" int impec_advect_x(impec_ctx *ctx, double dt) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    ctx->mob_valid = ctx->op_valid = false;
    TRY(copy_field(ctx, ctx->c, ctx->ta));
    TRY((sweep<0>(ctx, ctx->ta, ctx->c, nullptr, dt, 0, g.N1, -ng, g.N2 + ng)));
    // Slabs: the x sweep leaves the halo rows with the neighbours' values from BEFORE their sweep, where the undivided array holds the swept
    // rows; a sweep that follows (Strang: x, y, x) reads them.  The y sweep needs no such exchange: it sweeps the halo rows as well and
    // reproduces the neighbours' values bit for bit.
    return exchange_rows(ctx, ctx->c, g, ng);
}
extern "C" int impec_advect_y(impec_ctx *ctx, double dt) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    ctx->mob_valid = ctx->op_valid = false;
    TRY(copy_field(ctx, ctx->c, ctx->ta));
    return sweep<1>(ctx, ctx->ta, ctx->c, nullptr, dt, -ng, g.N1 + ng, 0, g.N2);
}

// CalculateIntercellFlux_F / _G (src/HyperbolicSplitSolver.h:19-22): the flux field of the CURRENT c and velocity for a step dt, in the
// reference's array shapes -- dir 0: F[(N1+1) x (N2+2g)], F(i, j) at out[i * (N2+2g) + (j+g)]; dir 1: G[(N1+2g) x (N2+1)], G(i, j) at
// out[(i+g) * (N2+1) + j].  c must carry filled ghosts (as for impec_advect_x / _y).  Slabs: every rank fills its own rows of the array.
extern "C" int impec_intercell_flux(impec_ctx *ctx, int dir, double dt, double *out_host) {
    if (!out_host || (dir != 0 && dir != 1)) return fail(ctx, IMPEC_ERR_ARG, "impec_intercell_flux: dir must be 0 (F) or 1 (G)");
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    const double phi = ctx->P.porosity; const bool phi1 = (phi == 1.0), up = ctx->P.hyper_method == IMPEC_UPWIND;
    const int i0 = dir == 0 ? 0 : -ng, i1 = dir == 0 ? g.N1 + 1 : g.N1 + ng, j0 = dir == 0 ? -ng : 0, j1 = dir == 0 ? g.N2 + ng : g.N2 + 1;
    const dim3 gr = grid2d(j1 - j0, i1 - i0);
    double *F = ctx->ta;
    const double *vel = dir == 0 ? ctx->u : ctx->v; const double hstep = dir == 0 ? ctx->D.dx : ctx->D.dy;
#define FLX(DR, M, PH) do { auto kern = k_intercell_flux<DR, M, PH>; LAUNCH(ctx, kern, gr, kBlock, g, (const double *)ctx->c, vel, F, dt, hstep, phi, i0, i1, j0, j1); } while (0)
    if (dir == 0) { if (up) { if (phi1) FLX(0, 1, true); else FLX(0, 1, false); } else { if (phi1) FLX(0, 0, true); else FLX(0, 0, false); } }
    else          { if (up) { if (phi1) FLX(1, 1, true); else FLX(1, 1, false); } else { if (phi1) FLX(1, 0, true); else FLX(1, 0, false); } }
#undef FLX
    const size_t w = (size_t)(j1 - j0) * sizeof(double);
    out_host += (size_t)g.i_off * (size_t)(j1 - j0);          // slabs: this rank's rows of the caller's whole-grid array (halo rows included: neighbours agree on them)
    CUDA_TRY(ctx, cudaMemcpy2DAsync(out_host, w, F + gidx(g, i0, j0), (size_t)g.ld * sizeof(double), w, (size_t)(i1 - i0), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bytes_d2h += (unsigned long long)w * (unsigned long long)(i1 - i0);
    return IMPEC_OK;
}

// src/HyperbolicSplitSolver.cpp:51-106
extern "C" int impec_advect_split(impec_ctx *ctx, double dt) {
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    ctx->mob_valid = ctx->op_valid = false;
    TRY(impec_fill_boundaries(ctx));
    switch (ctx->P.dim_splitting) {
    case IMPEC_ALTERNATING:
        if (ctx->prev_first_sweep == 1) { TRY(impec_advect_x(ctx, dt)); TRY(impec_advect_y(ctx, dt)); ctx->prev_first_sweep = 0; }
        else { TRY(impec_advect_y(ctx, dt)); TRY(impec_advect_x(ctx, dt)); ctx->prev_first_sweep = 1; }
        break;
    case IMPEC_X_FIRST: TRY(impec_advect_x(ctx, dt)); TRY(impec_advect_y(ctx, dt)); ctx->prev_first_sweep = 0; break;
    case IMPEC_Y_FIRST: TRY(impec_advect_y(ctx, dt)); TRY(impec_advect_x(ctx, dt)); ctx->prev_first_sweep = 1; break;
    case IMPEC_STRANG:  TRY(impec_advect_x(ctx, 0.5 * dt)); TRY(impec_advect_y(ctx, dt)); TRY(impec_advect_x(ctx, 0.5 * dt)); ctx->prev_first_sweep = 0; break;
    default:
        // c <- 1/2 ( Y(X(c)) + X(Y(c)) ); ghost strips are swept only where the second sweep reads them
        TRY(sweep<0>(ctx, ctx->c, ctx->ta, nullptr, dt, 0, g.N1, -ng, g.N2 + ng));        // X(c)      -> ta
        TRY(sweep<1>(ctx, ctx->ta, ctx->tb, nullptr, dt, 0, g.N1, 0, g.N2));              // Y(X(c))   -> tb
        TRY(sweep<1>(ctx, ctx->c, ctx->ta, nullptr, dt, -ng, g.N1 + ng, 0, g.N2));        // Y(c)      -> ta
        TRY(sweep<0>(ctx, ctx->ta, ctx->c, ctx->tb, dt, 0, g.N1, 0, g.N2));               // 1/2(tb + X(Y(c))) -> c
        ctx->prev_first_sweep = 0;
        break;
    }
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ a19 dispersion
// Crank-Nicolson half step (I - dt/2 L_D) c+ = (I + dt/2 L_D) c with the velocity-dependent dispersion tensor of
// src/MGSolver.cpp:846-919.  With alpha_l == alpha_t (all shipped cfgs) d12 = d21 = 0 and the operator is the 5-point
// links operator AW = dt/(2dx^2) d11, AS = dt/(2dy^2) d22 with mass term 1; Neumann folding (:1031-1289) again just
// removes the wall links.  Solved by the same PCG with one symmetric red-black sweep as preconditioner (the system
// is strongly diagonally dominant: dt D / dx^2 <~ cfl).  c must carry filled ghosts (the caller's FillBoundaries) and
// u, v their ghost fill (CalcVelocities).  x0 = 0 as upstream (:1379-1391).
// alpha_l != alpha_t: the Crank-Nicolson matrix is a full, NON-symmetric 9-point operator (src/MGSolver.cpp:923-948); upstream hands it to
// HYPRE's PCG regardless.  Here: matrix-free operator on the four face arrays of the dispersion tensor (k_cn9_*), solved by BiCGStab with
// the folded diagonal as preconditioner to the same relative residual (para_eps), x0 = 0.  Same linear system as the reference's (the oracle
// and the compiled reference agree on its entries bit for bit, tests/test_oracle_vs_ref.py); a general-purpose path, not a tuned one:
// no shipped cfg uses it.  Slabs: one halo row of the operand before every operator pass.
static int diffusion_split_general(impec_ctx *ctx, double dt, impec_solve_info *info) {
    const Grid &g = ctx->g;
    if (ctx->P.stype == IMPEC_HORIZONTAL) return fail(ctx, IMPEC_ERR_UNSUPPORTED, "dispersion with alpha_l != alpha_t is not combined with source_type horizontal");
    const dim3 gr = grid2d(g.N2, g.N1);
    const long long nparts = (long long)gr.x * gr.y;
    Level &L0 = ctx->L[0];
    ctx->op_valid = false; ctx->mob_valid = false;          // level-0 link arrays and the mobility scratch are reused
    Cn9 A; A.F11 = L0.AW; A.F12 = L0.AV; A.G22 = L0.AS; A.G21 = ctx->tb; A.dt = dt;
    LAUNCH(ctx, k_cn9_faces, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, (const double *)ctx->u, (const double *)ctx->v, ctx->D, L0.AW, L0.AV, L0.AS, ctx->tb);
    double *x = ctx->ta, *r = ctx->r, *r0 = ctx->z, *p = ctx->d, *v = ctx->s, *s = ctx->d2buf, *t = ctx->t0, *ph = ctx->MU, *sh = ctx->KV11, *dinv = L0.DINV;
    LAUNCH_COOP(ctx, k_cn9_rhs, gr, kBlock, g, A, ctx->D, (const double *)ctx->c, ctx->rhs, x, ctx->part);
    TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_BIPROD));
    TRY(fetch_scalars(ctx, S_BIPROD, 1));
    const double bb = ctx->h_scal[S_BIPROD], tol = ctx->P.para_eps;
    const size_t bytes = grid_elems(g) * sizeof(double);
    CUDA_TRY(ctx, cudaMemcpyAsync(r, ctx->rhs, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(r0, ctx->rhs, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(p, 0, bytes, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(v, 0, bytes, ctx->stream));
    // the folded diagonal (one operator pass on any ghost-filled field; its product is discarded)
    TRY(fill_scalar(ctx, r, 0));
    LAUNCH(ctx, k_cn9_apply, gr, kBlock, g, A, ctx->D, (const double *)r, t, dinv);
    auto dotp = [&](const double *a, const double *b, double *out) -> int {
        LAUNCH_COOP(ctx, k_dot, gr, kBlock, g, a, b, ctx->part);
        TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_TMP));
        TRY(fetch_scalars(ctx, S_TMP, 1));
        *out = ctx->h_scal[S_TMP];
        return IMPEC_OK;
    };
    double rho = 1., alpha = 1., omega = 1., rr = bb;
    int it = 0, converged = (bb > 0. && rr / bb >= tol * tol) ? 0 : 1;
    while (!converged && it < ctx->P.max_iter) {
        it++;
        double rho1; TRY(dotp(r0, r, &rho1));
        if (rho1 == 0.) break;
        const double beta = (rho1 / rho) * (alpha / omega);
        rho = rho1;
        LAUNCH(ctx, k_bicg_p, gr, kBlock, g, beta, omega, (const double *)r, (const double *)v, (const double *)dinv, p, ph);
        TRY(fill_scalar(ctx, ph, 0)); TRY(exchange_rows(ctx, ph, g, 1));        // mirrored ghosts at the walls, the neighbours' rows at slab edges
        LAUNCH(ctx, k_cn9_apply, gr, kBlock, g, A, ctx->D, (const double *)ph, v, (double *)nullptr);
        double r0v; TRY(dotp(r0, v, &r0v));
        alpha = rho / r0v;
        LAUNCH(ctx, k_bicg_s, gr, kBlock, g, alpha, (const double *)r, (const double *)v, (const double *)dinv, s, sh);
        TRY(fill_scalar(ctx, sh, 0)); TRY(exchange_rows(ctx, sh, g, 1));        // mirrored ghosts at the walls, the neighbours' rows at slab edges
        LAUNCH(ctx, k_cn9_apply, gr, kBlock, g, A, ctx->D, (const double *)sh, t, (double *)nullptr);
        double tt, ts; TRY(dotp(t, t, &tt)); TRY(dotp(t, s, &ts));
        omega = tt > 0. ? ts / tt : 0.;
        LAUNCH_COOP(ctx, k_bicg_xr, gr, kBlock, g, alpha, omega, (const double *)ph, (const double *)sh, (const double *)s, (const double *)t, x, r, ctx->part);
        TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_IPROD));
        TRY(fetch_scalars(ctx, S_IPROD, 1));
        rr = ctx->h_scal[S_IPROD];
        if (!(rr == rr)) return fail(ctx, IMPEC_ERR_NOCONV, "BiCGStab produced NaN at iteration %d", it);
        if (rr / bb < tol * tol) { converged = 1; break; }
        if (omega == 0.) break;
    }
    if (info) { info->iterations = it; info->converged = converged; info->relres = bb > 0. ? sqrt(rr / bb) : 0.; info->true_relres = info->relres; }
    LAUNCH(ctx, k_copy_interior, gr, kBlock, g, (const double *)x, ctx->c);
    // the face arrays carry values on the wall faces and in the pad entries, which the link kernels expect to be zero (wall links are
    // never written by the operator build): hand the arrays back clean
    CUDA_TRY(ctx, cudaMemsetAsync(L0.AW, 0, bytes, ctx->stream)); CUDA_TRY(ctx, cudaMemsetAsync(L0.AS, 0, bytes, ctx->stream)); CUDA_TRY(ctx, cudaMemsetAsync(L0.AV, 0, bytes, ctx->stream));
    return converged ? IMPEC_OK : fail(ctx, IMPEC_WARN_NOCONV, "BiCGStab did not reach eps=%g in %d iterations (relres %g)", tol, it, sqrt(rr / bb));
}

extern "C" int impec_diffusion_split(impec_ctx *ctx, double dt, impec_solve_info *info) {
    if (ctx->P.alpha_l != ctx->P.alpha_t) return diffusion_split_general(ctx, dt, info);
    const Grid &g = ctx->g;
    const dim3 gr = grid2d(g.N2, g.N1);
    const long long nparts = (long long)gr.x * gr.y;
    Level &L0 = ctx->L[0];
    ctx->op_valid = false; ctx->mob_valid = false;          // the level-0 link arrays are reused for the parabolic operator
    LAUNCH(ctx, k_disp_links, grid2d(g.N2 + 1, g.N1 + 1), kBlock, g, (const double *)ctx->u, (const double *)ctx->v, dt, ctx->D, L0.AW, L0.AS, L0.AV);
    if (L0.dist) { TRY(exchange_rows(ctx, L0.AW, L0.g, IMPEC_GH)); TRY(exchange_rows(ctx, L0.AS, L0.g, IMPEC_GH)); TRY(exchange_rows(ctx, L0.AV, L0.g, IMPEC_GH)); }
    LevelDev A = level_dev(ctx, 0, 1.0);
    A.pin = A.pin_end = -1; A.sig_idx = A.sig_end = -1;
    if (ctx->P.stype == IMPEC_HORIZONTAL && g.wall_hi) {
        // Dirichlet row i = N1-2 of the Crank-Nicolson problem (src/MGSolver.cpp:130-139, 1292-1315): c = 0 there (rhs zeroed, :1357-1364)
        A.pin = (long long)gidx(g, g.N1 - 2, 0); A.pin_end = A.pin + g.N2;
    }
    TRY(finish_level(ctx, 0, A, nullptr));
    Cycle C; C.ctx = ctx; C.D.assign(1, A); C.ncoarse = 1; C.w_from = 1; C.w_to = 0; C.nl = 1;
    double *x = ctx->ta;
    LAUNCH_COOP(ctx, k_cn_rhs, gr, kBlock, A, (const double *)ctx->c, ctx->rhs, x, ctx->part);
    TRY(reduce_final(ctx, RED_SUM, nparts, 1, 0, S_BIPROD));
    const int rc = pcg_solve(ctx, C, ctx->rhs, x, ctx->P.para_eps, info, true);      // k_cn_rhs zeroed x (x0 = 0 as upstream)
    if (rc < 0) return rc;                                                      // IMPEC_WARN_NOCONV: keep the last iterate as upstream does
    LAUNCH(ctx, k_copy_interior, gr, kBlock, g, (const double *)x, ctx->c);     // MapHYPREVectorToVertex: interior only
    return rc;
}

// ------------------------------------------------------------------------------------------------ 8f-3 projection
// Impec2DSolver::Project (src/Impec2DSolver.cpp:690-716): the concentration (field 0) or the pressure (field 1) of a PARALLEL
// run interpolated onto the diagonal grid of round(N/sqrt 2) cells; out_host receives dN1 x dN2 doubles (row-major, no ghosts).
extern "C" int impec_project_dims(const impec_ctx *ctx, int *dN1, int *dN2) {
    if (dN1) *dN1 = (int)round(ctx->P.N1 / sqrt(2.)); if (dN2) *dN2 = (int)round(ctx->P.N2 / sqrt(2.));
    return IMPEC_OK;
}
extern "C" int impec_project_to_diagonal(impec_ctx *ctx, int field, double *out_host) {
    if (field != 0 && field != 1) return fail(ctx, IMPEC_ERR_ARG, "field must be 0 (c) or 1 (p)");
    if (!out_host) return fail(ctx, IMPEC_ERR_ARG, "null output");
    const impec_params &P = ctx->P; const Grid &g = ctx->g;
    ProjectArgs A;
    impec_project_dims(ctx, &A.dN1, &A.dN2);
    const double angle = -M_PI / 4.;
    A.x_lb = P.x_lb; A.y_lb = P.y_lb; A.x_ub = P.x_ub; A.y_ub = P.y_ub; A.ca = cos(angle); A.sa = sin(angle);
    A.p_dx = (P.x_ub - P.x_lb) / P.N1; A.p_dy = (P.y_ub - P.y_lb) / P.N2;
    A.d_dx = (P.x_ub - P.x_lb) / (sqrt(2.) * A.dN1); A.d_dy = (P.y_ub - P.y_lb) / (sqrt(2.) * A.dN2);
    A.ng = P.g;
    // Slabs (collective; every rank receives the whole projected field): the diagonal grid is processed in chunks of rows that fit the
    // slab-sized scratch field; each rank evaluates the cells whose stencil starts in its rows, the chunk is summed over the ranks.
    A.N1g = ctx->N1g; A.i_off = g.i_off;
    A.own_lo = g.wall_lo ? INT_MIN : g.i_off; A.own_hi = g.wall_hi ? INT_MAX : g.i_off + g.N1;
    const long long chunk = std::min<long long>(A.dN1, (long long)grid_elems(g) / A.dN2);
    if (chunk < 1) return fail(ctx, IMPEC_ERR_ARG, "diagonal grid row longer than the scratch field");
    double *dst = ctx->tb;                                   // transport scratch (free between steps)
    double *src = field == 0 ? ctx->c : ctx->p;
    TRY(exchange_rows(ctx, src, g, 1));                       // the stencil of a cell owned through its lower row reaches one row into the next slab
    for (long long r0 = 0; r0 < A.dN1; r0 += chunk) {
        A.row0 = (int)r0; A.nrows = (int)std::min<long long>(chunk, A.dN1 - r0);
        const size_t n = (size_t)A.nrows * A.dN2;
        LAUNCH(ctx, k_project_parallel, grid2d(A.dN2, A.nrows), kBlock, g, (const double *)src, A, dst);
        if (ctx->nranks > 1) TRY(comm_allreduce(ctx, dst, (int)n, RED_SUM));
        CUDA_TRY(ctx, cudaMemcpyAsync(out_host + (size_t)r0 * A.dN2, dst, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->bytes_d2h += n * sizeof(double);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ 8f-2 binary checkpoint
// The reference restarts from a text file of c only (init_file, src/Impec2DSolver.cpp:194-220), which loses the pressure, the
// produced mass and the sweep parity and is unusable at 16384^2 (268 M text lines).  A checkpoint here is the exact device state
// of this rank: header + the padded c, p, u, v arrays (ghost / halo rows included) + produced mass + sweep parity, so a
// restarted run continues BIT FOR BIT.  Multi-GPU: one file per rank (the caller chooses the names).
// Version 2: little-endian marker and an FNV-1a hash of the physics in impec_params (a checkpoint must not be continued with other
// viscosities, bounds, ... by accident); written to "<path>.tmp" and renamed into place, so a crash during a save never destroys
// the previous restart point.
struct CkptHeader { char magic[8]; int version, N1, N2, g, rank, nranks, ld, rows; long long step; double t, outflux; int prev_first_sweep; unsigned endian;
                    unsigned long long params_hash; };
static unsigned long long ckpt_params_hash(const impec_params &P) {
    unsigned long long h = 1469598103934665603ULL;
    auto mix = [&](const void *p, size_t n) { const unsigned char *b = (const unsigned char *)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; } };
    const double d[] = {P.x_lb, P.x_ub, P.y_lb, P.y_ub, P.mu_res, P.mu_inj, P.nu, P.porosity, P.inj_fluid_conc, P.alpha_l, P.alpha_t};
    const int i[] = {P.N1, P.N2, P.g, P.avgproc, P.is_test, P.testno, P.stype, P.diffusion, P.dim_splitting};
    mix(d, sizeof(d)); mix(i, sizeof(i));
    return h;
}
extern "C" int impec_checkpoint_save(impec_ctx *ctx, const char *path, double t, long long step) {
    if (!path) return fail(ctx, IMPEC_ERR_ARG, "null path");
    const Grid &g = ctx->g;
    CkptHeader H; memset(&H, 0, sizeof(H));
    memcpy(H.magic, "IMPECCK1", 8); H.version = 2; H.endian = 0x01020304u; H.params_hash = ckpt_params_hash(ctx->P); H.N1 = ctx->P.N1; H.N2 = ctx->P.N2; H.g = ctx->P.g; H.rank = ctx->rank; H.nranks = ctx->nranks;
    H.ld = g.ld; H.rows = g.rows; H.step = step; H.t = t; H.prev_first_sweep = ctx->prev_first_sweep;
    CUDA_TRY(ctx, cudaMemcpyAsync(&H.outflux, ctx->scal + S_OUTFLUX, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(ctx, IMPEC_ERR_ARG, "cannot open %s for writing", tmp.c_str());
    bool ok = fwrite(&H, sizeof(H), 1, f) == 1;
    const size_t n = grid_elems(g);
    std::vector<double> buf(n);
    for (const double *fld : {(const double *)ctx->c, (const double *)ctx->p, (const double *)ctx->u, (const double *)ctx->v}) {
        if (cudaMemcpyAsync(buf.data(), fld, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ok = false; break; }
        ok = ok && fwrite(buf.data(), sizeof(double), n, f) == n;
    }
    ok = (fflush(f) == 0) && ok;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); return fail(ctx, IMPEC_ERR_CUDA, "writing checkpoint %s failed", path); }
    if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(ctx, IMPEC_ERR_ARG, "cannot move %s into place", path); }
    return IMPEC_OK;
}
extern "C" int impec_checkpoint_load(impec_ctx *ctx, const char *path, double *t, long long *step) {
    if (!path) return fail(ctx, IMPEC_ERR_ARG, "null path");
    const Grid &g = ctx->g;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(ctx, IMPEC_ERR_ARG, "cannot open %s", path);
    CkptHeader H;
    if (fread(&H, sizeof(H), 1, f) != 1 || memcmp(H.magic, "IMPECCK1", 8) != 0 || H.version != 2 || H.endian != 0x01020304u) { fclose(f); return fail(ctx, IMPEC_ERR_ARG, "%s is not an impec2d_b200 (version 2, same byte order) checkpoint", path); }
    if (H.params_hash != ckpt_params_hash(ctx->P)) { fclose(f); return fail(ctx, IMPEC_ERR_ARG, "checkpoint %s was written with different physical parameters (bounds, viscosities, nu, porosity, ...)", path); }
    if (H.N1 != ctx->P.N1 || H.N2 != ctx->P.N2 || H.g != ctx->P.g || H.rank != ctx->rank || H.nranks != ctx->nranks || H.ld != g.ld || H.rows != g.rows) {
        fclose(f);
        return fail(ctx, IMPEC_ERR_ARG, "checkpoint %s is for a %d x %d grid on rank %d of %d; this context is %d x %d on rank %d of %d", path, H.N1, H.N2, H.rank, H.nranks,
                    ctx->P.N1, ctx->P.N2, ctx->rank, ctx->nranks);
    }
    const size_t n = grid_elems(g);
    std::vector<double> buf(n);
    bool ok = true;
    for (double *fld : {ctx->c, ctx->p, ctx->u, ctx->v}) {
        ok = ok && fread(buf.data(), sizeof(double), n, f) == n;
        if (!ok) break;
        if (cudaMemcpyAsync(fld, buf.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ok = false; break; }
    }
    fclose(f);
    if (!ok) return fail(ctx, IMPEC_ERR_ARG, "checkpoint %s is truncated or unreadable", path);
    LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->scal + S_OUTFLUX, H.outflux);
    ctx->prev_first_sweep = H.prev_first_sweep;
    ctx->mob_valid = ctx->op_valid = false; ctx->dt_cached = false;
    if (t) *t = H.t; if (step) *step = H.step;
    return IMPEC_OK;
}

// ------------------------------------------------------------------------------------------------ a14 transport step
// A solver that stops at max_iter (IMPEC_WARN_NOCONV) does not abort the step: the step is finished with the last iterate, as
// upstream, and the warning is returned at the end.
#define TRY_WARN(call, warn) do { int rc_ = (call); if (rc_ < 0) return rc_; if (rc_ > 0) (warn) = rc_; } while (0)
extern "C" int impec_transport_step(impec_ctx *ctx, double dt) {
    int warn = IMPEC_OK;
    TRY(impec_source_split(ctx, 0.5 * dt));
    TRY(impec_fill_boundaries(ctx));
    if (ctx->P.diffusion) TRY_WARN(impec_diffusion_split(ctx, 0.5 * dt, &ctx->last_para), warn);
    TRY(impec_advect_split(ctx, dt));
    TRY(impec_source_split(ctx, 0.5 * dt));
    TRY(impec_fill_boundaries(ctx));
    if (ctx->P.diffusion) TRY_WARN(impec_diffusion_split(ctx, 0.5 * dt, &ctx->last_para), warn);
    return warn;
}

// ------------------------------------------------------------------------------------------------ a20 IMPES step
extern "C" int impec_impes_step(impec_ctx *ctx, double t, double t_next_output, double *dt_used, impec_solve_info *info) {
    int warn = IMPEC_OK;
    TRY(impec_fill_boundaries(ctx));
    TRY_WARN(impec_pressure_solve(ctx, ctx->P.eps, info), warn);
    TRY(impec_calc_velocities(ctx));
    double dt = 0.;
    TRY(impec_max_stable_dt(ctx, &dt));
    if (t_next_output > t && t + dt >= t_next_output) dt = t_next_output - t;      // src/Impec2DSolver.cpp:375-378
    TRY_WARN(impec_transport_step(ctx, dt), warn);
    if (dt_used) *dt_used = dt;
    return warn;
}

extern "C" int impec_diagnostics(impec_ctx *ctx, impec_diag *out) {
    const Grid &g = ctx->g;
    const dim3 gr = grid2d(g.N2, g.N1);
    const long long nparts = (long long)gr.x * gr.y;
    LAUNCH_COOP(ctx, k_diag_cells, gr, kBlock, g, ctx->c, ctx->u, ctx->v, ctx->q, ctx->D.dx, ctx->D.dy, ctx->part, ctx->part_cap);
    TRY(reduce_final(ctx, RED_SUM, nparts, 2, ctx->part_cap, S_DIAG));
    LAUNCH_COOP(ctx, k_reduce_final<RED_MAX>, dim3(1), dim3(1024), ctx->part + 2 * ctx->part_cap, nparts, 0LL, ctx->scal + S_DIAG + 2);
    LAUNCH_COOP(ctx, k_reduce_final<RED_MIN>, dim3(1), dim3(1024), ctx->part + 3 * ctx->part_cap, nparts, 0LL, ctx->scal + S_DIAG + 3);
    TRY(comm_allreduce(ctx, ctx->scal + S_DIAG + 2, 1, RED_MAX));
    TRY(comm_allreduce(ctx, ctx->scal + S_DIAG + 3, 1, RED_MIN));
    const int nb = (std::max(g.N1, g.N2) + 255) / 256;
    LAUNCH_COOP(ctx, k_diag_boundary, dim3(nb), dim3(256), g, ctx->u, ctx->v, ctx->part);
    TRY(reduce_final(ctx, RED_SUM, nb, 1, 0, S_DIAG + 4));
    TRY(fetch_scalars(ctx, S_OUTFLUX, S_DIAG + 5 - S_OUTFLUX));
    out->total_mass = ctx->h_scal[S_DIAG];
    out->divergence_error = ctx->h_scal[S_DIAG + 1] * ctx->D.dx * ctx->D.dx;       // dx*dx as upstream (:556)
    out->c_max = ctx->h_scal[S_DIAG + 2]; out->c_min = ctx->h_scal[S_DIAG + 3];
    out->boundary_velocity_error = ctx->h_scal[S_DIAG + 4];
    out->total_mass_outflux = ctx->h_scal[S_OUTFLUX];
    return IMPEC_OK;
}

// the integral diagnostics the north-star compares fingering runs on, as device reductions (the host-side
// definitions live in impec2d_b200/diagnostics.py and are checked against these on downloaded fields in the tests).
extern "C" int impec_integral_diagnostics(impec_ctx *ctx, double lo, double hi, impec_integral_diag *out) {
    if (!out) return fail(ctx, IMPEC_ERR_ARG, "null output");
    if (!ctx->have_q) return fail(ctx, IMPEC_ERR_ARG, "integral diagnostics before impec_upload_q");
    const Grid &g = ctx->g;
    const dim3 gr = grid2d(g.N2, g.N1);
    const long long nparts = (long long)gr.x * gr.y;
    LAUNCH_COOP(ctx, k_diag_mixing, gr, kBlock, g, (const double *)ctx->c, lo, hi, ctx->part, ctx->part_cap);
    TRY(reduce_final(ctx, RED_SUM, nparts, 2, ctx->part_cap, S_DIAG + 5));          // S_DIAG+5: cells in the mixing zone, +6: sum c
    if (ctx->nwell) {
        const int nb = (ctx->nwell + 255) / 256;
        LAUNCH_COOP(ctx, k_diag_wells, dim3((unsigned)nb), dim3(256), ctx->nwell, ctx->well_i, ctx->well_j, ctx->well_q, g, (const double *)ctx->c, ctx->part, ctx->part_cap);
        LAUNCH_COOP(ctx, k_reduce_final<RED_SUM>, dim3(2), dim3(1024), ctx->part, (long long)nb, ctx->part_cap, ctx->scal + S_DIAG + 7);
    } else {
        LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->scal + S_DIAG + 7, 0.0);
        LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->scal + S_DIAG + 8, 0.0);
    }
    TRY(comm_allreduce(ctx, ctx->scal + S_DIAG + 7, 2, RED_SUM));
    TRY(fetch_scalars(ctx, S_DIAG + 5, 4));
    const double lx = ctx->P.x_ub - ctx->P.x_lb, ly = ctx->P.y_ub - ctx->P.y_lb;
    out->mixing_zone_cells = ctx->h_scal[S_DIAG + 5];
    out->mixing_zone_length = ctx->h_scal[S_DIAG + 5] * ctx->D.dx * ctx->D.dy / sqrt(lx * lx + ly * ly);
    out->mass_in_place = ctx->h_scal[S_DIAG + 6] * ctx->D.dx * ctx->D.dy;
    out->produced_fraction = ctx->h_scal[S_DIAG + 7] > 0. ? ctx->h_scal[S_DIAG + 8] / ctx->h_scal[S_DIAG + 7] : 0.;
    return IMPEC_OK;
}

#include "host_util.inl"

// ------------------------------------------------------------------------------------------------ profiling hook
// Launches one named piece of the path `reps` times on the context stream between two events and returns the
// average duration.  Used by bench.py for the per-kernel roofline numbers (never under a profiler).
extern "C" int impec_profile_kernel(impec_ctx *ctx, const char *name, int reps, double *avg_ms) {
    if (!name || reps < 1 || !avg_ms) return fail(ctx, IMPEC_ERR_ARG, "bad arguments to impec_profile_kernel");
    const std::string k(name);
    const Grid &g = ctx->g; const int ng = ctx->P.g;
    if (!ctx->op_valid) TRY(impec_build_operator(ctx));
    Cycle C; TRY(make_cycle(ctx, C, 0.0, 0));
    const dim3 gr = grid2d(g.N2, g.N1);
    Level &L0 = ctx->L[0];
    auto body = [&]() -> int {
        // "<kernel>@<level>": the tile kernels on a coarser level (replicated or single-GPU levels)
        const size_t at = k.find('@');
        if (at != std::string::npos) {
            const int l = atoi(k.c_str() + at + 1);
            const std::string kk = k.substr(0, at);
            if (l < 0 || l + 1 >= (int)ctx->L.size()) return fail(ctx, IMPEC_ERR_ARG, "no such level: %s", k.c_str());
            Level &L = ctx->L[l];
            const double *fl = l == 0 ? ctx->r : L.f;
            if (kk == "fused_pre") return fused_pre(C, l, fl);
            if (kk == "fused_post") return fused_post(C, l, fl, nullptr, nullptr);
            if (kk == "fused_plain") return fused_plain(C, l, fl, true);
            if (kk == "cycle") return mg_cycle(C, l, fl);
            return fail(ctx, IMPEC_ERR_ARG, "unknown kernel name '%s'", k.c_str());
        }
        if (k == "fused_plain") return fused_plain(C, 0, ctx->r, true);
        if (k == "fused_pre") return fused_pre(C, 0, ctx->r);
        if (k == "fused_post") return fused_post(C, 0, ctx->r, ctx->part, nullptr);
        if (k == "smooth_sweep") return smooth(C, 0, ctx->r, L0.x, true, false);
        if (k == "smooth_sweep_zero") return smooth(C, 0, ctx->r, L0.x, true, true);
        if (k == "residual") { LAUNCH(ctx, k_residual, gr, kBlock, C.D[0], ctx->r, (const double *)L0.x, L0.r); return IMPEC_OK; }
        if (k == "restrict") { Level &Lc = ctx->L[1]; LAUNCH(ctx, k_restrict, grid2d(Lc.g.N2, Lc.g.N1), kBlock, L0.g, L0.r, Lc.g, Lc.f); return IMPEC_OK; }
        if (k == "prolong_add") { Level &Lc = ctx->L[1]; LAUNCH(ctx, k_prolong_add, gr, kBlock, Lc.g, Lc.x, L0.g, L0.x, L0.pin, L0.pin_end); return IMPEC_OK; }
        if (k == "pcg_matvec" && ctx->matvec_march) { LAUNCH_COOP(ctx, k_pcg_matvec_march<false>, matvec_march_grid(g), dim3(MV_BLOCK), chunked(C.D[0], 0, 1, ctx->pf_mv), (const double *)(ctx->scal + S_TMP), (const double *)L0.x, (const double *)ctx->d, ctx->d2buf, ctx->s, ctx->part, matvec_march_rows(g)); return IMPEC_OK; }
        if (k == "pcg_matvec") { LAUNCH_COOP_SMEM(ctx, k_pcg_matvec<false>, fused_grid<FUSED_PLAIN>(g), kFusedBlock, kFusedSmem, C.D[0], (const double *)(ctx->scal + S_TMP), (const double *)L0.x, (const double *)ctx->d, ctx->d2buf, ctx->s, ctx->part); return IMPEC_OK; }
        if (k == "matvec_dot") { LAUNCH_COOP(ctx, k_matvec_dot, gr, kBlock, C.D[0], ctx->d, ctx->s, ctx->part); return IMPEC_OK; }
        if (k == "update_xr") { LAUNCH_COOP(ctx, k_update_xr<0>, update_xr_grid(g), dim3(UX_BLOCK), g, ctx->scal + S_TMP, ctx->d, ctx->s, ctx->t0, ctx->rhs, ctx->part, (const double *)nullptr, (double *)nullptr); return IMPEC_OK; }
        if (k == "update_r") { LAUNCH_COOP(ctx, k_update_xr<1>, update_xr_grid(g), dim3(UX_BLOCK), g, ctx->scal + S_TMP, ctx->d, ctx->s, ctx->t0, ctx->rhs, ctx->part, (const double *)nullptr, ctx->scal + S_ALPHA_PREV); return IMPEC_OK; }
        if (k == "update_xxr") { LAUNCH_COOP(ctx, k_update_xr<2>, update_xr_grid(g), dim3(UX_BLOCK), g, ctx->scal + S_TMP, ctx->d, ctx->s, ctx->t0, ctx->rhs, ctx->part, (const double *)ctx->d2buf, ctx->scal + S_ALPHA_PREV); return IMPEC_OK; }
        if (k == "dot") { LAUNCH_COOP(ctx, k_dot, gr, kBlock, g, ctx->r, ctx->z, ctx->part); return IMPEC_OK; }
        if (k == "update_d") { LAUNCH(ctx, k_update_d<false>, gr, kBlock, g, ctx->scal + S_TMP, ctx->z, ctx->d); return IMPEC_OK; }
        if (k == "vcycle") return mg_cycle(C, 0, ctx->r);
        if (k == "tail") { const int tl = ctx->tail_level >= 0 ? ctx->tail_level : (int)ctx->L.size() - 1; return mg_cycle(C, tl, ctx->L[tl].f); }
        // communication primitives as the solver issues them (multi-rank contexts; no-ops on one rank)
        if (k == "halo_x2") return exchange_rows(ctx, L0.x, g, 2);
        if (k == "halo_f3") return exchange_rows(ctx, ctx->r, g, IMPEC_HALO);
        if (k == "allreduce1") return comm_allreduce(ctx, ctx->scal + S_TMP, 1, RED_MAX);
        if (k == "gather_rhs") { for (int l = 0; l + 1 < (int)ctx->L.size(); l++) if (at_transition(ctx, l)) return gather_coarse_rhs(ctx, l); return IMPEC_OK; }
        if (k == "sweep_x") return sweep<0>(ctx, ctx->c, ctx->ta, nullptr, 0.0, 0, g.N1, -ng, g.N2 + ng);
        if (k == "sweep_y") return sweep<1>(ctx, ctx->c, ctx->ta, nullptr, 0.0, -ng, g.N1 + ng, 0, g.N2);
        if (k == "transport_step") return impec_transport_step(ctx, ctx->profile_dt);
        if (k == "diffusion_split") return impec_diffusion_split(ctx, 0.5 * ctx->profile_dt, nullptr);
        if (k == "build_operator") return impec_build_operator(ctx);
        if (k == "calc_velocities") return impec_calc_velocities(ctx);
        return fail(ctx, IMPEC_ERR_ARG, "unknown kernel name '%s'", k.c_str());
    };
    if (k == "update_xr" || k == "update_r" || k == "update_xxr" || k == "update_d" || k == "pcg_matvec") {      // alpha = beta = 0/1 so the vectors stay finite
        LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->scal + S_TMP, 0.0);
        LAUNCH(ctx, k_set_scalar, dim3(1), dim3(1), ctx->scal + S_TMP + 1, 1.0);
    }
    if (k == "restrict" || k == "prolong_add") { if (ctx->L.size() < 2) return fail(ctx, IMPEC_ERR_ARG, "single-level hierarchy"); }
    TRY(body());                                   // warm-up
#ifndef IMPEC_EMU
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0)); CUDA_TRY(ctx, cudaEventCreate(&e1));
    CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < reps; i++) TRY(body());
    CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *avg_ms = (double)ms / reps;
#else
    for (int i = 0; i < reps; i++) TRY(body());
    *avg_ms = 0.0;
#endif
    return IMPEC_OK;
}
