// kernels_tail.inl -- the multigrid TAIL (all levels below ~2 M cells) as ONE persistent launch.
//
// The W-cycle visits the tail four times per cycle (its lower half eight times) and every visit used to be ~25 launches of
// 5-13 us each, replayed from a CUDA graph: 0.21 ms per visit, 0.87 ms per cycle on ANY number of GPUs (a third of the 8-GPU
// cycle, profiles/r1m_comm_diag_n8.json).  These kernels are not bandwidth-bound; their cost is launch gaps and the
// fill/drain of a handful of CTAs.  k_mg_tail runs the whole visit as a PROGRAM of phases interpreted by one resident CTA per
// SM: a phase is one of the tile kernels' bodies (PRE / POST / PLAIN / MID, 16-row tiles, looped over the level's tiles) or the
// single-CTA bottom-of-hierarchy V-cycle (small_body), and phases are separated by a grid-wide barrier (one atomic per CTA on
// a counter in global memory, system of release / acquire fences) instead of a kernel boundary.
//
// MEASURED (B200, 16384^2): 0.297 ms per visit against 0.214 ms for the graph replay, so the option (mg_persist_tail) is OFF by
// default: with one CTA per SM (the bottom V-cycle needs 125 KB of shared memory) every phase is a latency-bound loop over
// tiles, and a grid barrier costs about what a kernel boundary inside a CUDA graph costs.  The tail is bound by its chain of
// ~25 DEPENDENT phases, not by launch overhead.  The kernel stays as a tested option and as the record of that experiment.
//
// * The program is recorded on the host by the same recursion that issues the launches (tail_record in impec2d_b200.cu), so
//   the arithmetic and its order are those of the launch path: results are bit-identical.
// * Iterates and right-hand sides are rewritten by other CTAs between phases of the same launch and L1 is coherent only at
//   launch boundaries, so those arrays are read with ld.global.cg (template parameter CG of fused_body / small_body).
// * All CTAs must be co-resident: the kernel is launched with cudaLaunchCooperativeKernel on a grid of at most one CTA per SM.
//   A barrier that waits longer than TAIL_TIMEOUT_CYCLES raises a flag and gives up instead of hanging the GPU.
enum { TAIL_PRE = 0, TAIL_POST = 1, TAIL_PLAIN = 2, TAIL_MID = 3, TAIL_SMALL = 4 };
#define TAIL_TH 16
#define TAIL_TIMEOUT_CYCLES 4000000000LL      /* ~2 s */
struct TailPhase {
    int op, c0, nbx, nby;
    LevelDev D;
    Grid gc;
    const double *f, *xin, *ec;
    double *x, *fc;
    int small_idx;                            // TAIL_SMALL: index into TailProgram::small
};
#define TAIL_MAXPHASES 48
#define TAIL_MAXSMALL 2
struct TailProgram {
    int nphases;
    TailPhase ph[TAIL_MAXPHASES];
    SmallArgs small[TAIL_MAXSMALL];
};

IMPEC_D void tail_barrier(unsigned *ctr, unsigned &target, double *err) {
    IMPEC_SYNCTHREADS();
#ifndef IMPEC_EMU
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        if (err[0] == 0.) {
            const long long t0 = clock64();
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
                if (v >= target) break;
                if (clock64() - t0 > TAIL_TIMEOUT_CYCLES) { err[0] = 2.0; break; }
            } while (true);
        }
        __threadfence();
    }
    IMPEC_SYNCTHREADS();
#endif
}

template <int MODE, bool F32>
IMPEC_D void tail_tiles(const TailPhase &P, double *X) {
    constexpr int HI = FusedGeom<MODE, TAIL_TH>::HI, TI = FusedGeom<MODE, TAIL_TH>::TI;
    const int ntiles = P.nbx * P.nby;
    for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x) {
        const int bx = tile % P.nbx, by = tile / P.nbx;
        const int j0 = bx * FT_TJ - FT_HJ, i0 = by * TI - HI;
        bool edge = i0 < P.D.g.rlo || j0 < 0 || i0 + TAIL_TH >= P.D.g.rhi || j0 + FT_W >= P.D.g.N2;
        if (P.D.pin >= 0) {
            const int pi = (int)(P.D.pin / P.D.g.ld) - IMPEC_GH, pj = (int)(P.D.pin % P.D.g.ld) - IMPEC_PJ;
            edge = edge || (pi >= i0 && pi < i0 + TAIL_TH && pj >= j0 && pj < j0 + FT_W);
        }
        if (edge) fused_body<MODE, true, F32, TAIL_TH, true>(P.D, P.f, P.xin, P.x, P.c0, P.gc, P.ec, P.fc, (double *)nullptr, X, 1, bx, by, P.nbx);
        else fused_body<MODE, false, F32, TAIL_TH, true>(P.D, P.f, P.xin, P.x, P.c0, P.gc, P.ec, P.fc, (double *)nullptr, X, 1, bx, by, P.nbx);
        IMPEC_SYNCTHREADS();                   // the tile in shared memory is reused by the next tile of this CTA
    }
}

template <bool F32>
__global__ void __launch_bounds__(FT_W / 2 * FT_NTY, 1) k_mg_tail(const TailProgram *__restrict__ prog, unsigned *ctr, double *err) {
    IMPEC_DYN_SMEM(double, X);
    unsigned target = 0;
    const int np = prog->nphases;
    for (int p = 0; p < np; p++) {
        const TailPhase &P = prog->ph[p];
        switch (P.op) {
        case TAIL_PRE:   tail_tiles<FUSED_PRE, F32>(P, X); break;
        case TAIL_POST:  tail_tiles<FUSED_POST, F32>(P, X); break;
        case TAIL_PLAIN: tail_tiles<FUSED_PLAIN, F32>(P, X); break;
        case TAIL_MID:   tail_tiles<FUSED_MID, F32>(P, X); break;
        default:
            if (blockIdx.x == 0) small_body<F32, FT_W / 2 * FT_NTY, true>(prog->small[P.small_idx], P.f, X, (int)(threadIdx.y * blockDim.x + threadIdx.x));
            break;
        }
        if (p + 1 < np) tail_barrier(ctr, target, err);
    }
}
