// common.h -- device data layout, indexing and reduction helpers shared by all kernels.
#pragma once
#include "shim.h"

// Every 2-D field (cell-, face- or vertex-centred, on every multigrid level) lives in one padded
// layout: rows are x (i, slow), columns y (j, contiguous) exactly like the reference's blitz arrays
// (src/GenSimDefs.h:25-31), with GH ghost rows on either side and PJ pad columns on the left so that
// the first interior element of every row is 32-byte aligned and rows start on 128-byte boundaries.
#define IMPEC_GH 4      // allocated halo rows (>= the reference's num_ghostcells = 2; 4 so that the fused multigrid kernels
                        // can treat 3 halo rows of a neighbouring slab as active cells)
#define IMPEC_HALO 3    // halo rows of a neighbouring slab that carry valid cells
#define IMPEC_PJ 4      // left pad in doubles (>= ghost width, multiple of 4)

struct Grid {           // geometry of one level (or of the fine grid) as seen by one rank (a slab of rows)
    int N1, N2;         // interior cells of this slab
    int ld;             // pitch in doubles
    int rows;           // allocated rows = N1 + 1 + 2*GH
    int wall_lo, wall_hi;   // 1 if row 0 / row N1 is a physical wall of the domain, 0 if a neighbouring slab continues there
    int rlo, rhi;       // rows [rlo, rhi) hold valid cells (interior plus IMPEC_HALO rows towards neighbouring slabs)
    int i_off;          // global row index of local row 0
};

IMPEC_HD size_t gidx(const Grid &g, int i, int j) { return (size_t)(i + IMPEC_GH) * (size_t)g.ld + (size_t)(j + IMPEC_PJ); }
IMPEC_HD size_t grid_elems(const Grid &g) { return (size_t)g.rows * (size_t)g.ld; }

static inline Grid make_grid(int N1, int N2);
// slab of a distributed level: rank `rank` of `nranks` owns N1g/nranks rows
static inline Grid make_slab(int N1g, int N2, int rank, int nranks);
static inline Grid make_grid(int N1, int N2) {
    Grid g;
    g.N1 = N1; g.N2 = N2;
    g.ld = ((N2 + 1 + 2 * IMPEC_PJ) + 15) / 16 * 16;
    g.rows = N1 + 1 + 2 * IMPEC_GH;
    g.wall_lo = g.wall_hi = 1; g.rlo = 0; g.rhi = N1; g.i_off = 0;
    return g;
}

static inline Grid make_slab(int N1g, int N2, int rank, int nranks) {
    Grid g = make_grid(N1g / nranks, N2);
    g.i_off = rank * (N1g / nranks);
    g.wall_lo = rank == 0; g.wall_hi = rank == nranks - 1;
    g.rlo = g.wall_lo ? 0 : -IMPEC_HALO;
    g.rhi = g.N1 + (g.wall_hi ? 0 : IMPEC_HALO);
    return g;
}

// ---- 128-bit accesses of an aligned pair of doubles
struct d2 { double x, y; };
IMPEC_D d2 ld2(const double *p) {
#ifndef IMPEC_EMU
    const double2 v = *reinterpret_cast<const double2 *>(p);
    d2 r; r.x = v.x; r.y = v.y; return r;
#else
    d2 r; r.x = p[0]; r.y = p[1]; return r;
#endif
}
IMPEC_D void st2(double *p, double a, double b) {
#ifndef IMPEC_EMU
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}
// software prefetch of a line into L2 (no register is held; used by the marching kernels a few rows ahead of their loads)
IMPEC_D void pf_l2(const void *p) {
#ifndef IMPEC_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ---- reductions -------------------------------------------------------------------------------
// Stage 1: every block reduces its values with warp shuffles + one shared-memory hop and writes ONE
// partial to part[blockIdx]; stage 2 (k_reduce_final) combines the partials in a fixed pairwise tree,
// so results are deterministic and independent of scheduling.
enum { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

template <int OP> IMPEC_D double red_op(double a, double b) {
    if (OP == RED_SUM) return a + b;
    if (OP == RED_MAX) return a > b ? a : b;
    return a < b ? a : b;
}
template <int OP> IMPEC_D double red_identity() {
    if (OP == RED_SUM) return 0.0;
    if (OP == RED_MAX) return -DBL_MAX;
    return DBL_MAX;
}

template <int OP> IMPEC_D double warp_reduce(double v) {
    for (int d = 16; d > 0; d >>= 1) v = red_op<OP>(v, IMPEC_SHFL_DOWN(v, d));
    return v;
}

// all threads of the block must call this; result valid in thread 0
template <int OP> IMPEC_D double block_reduce(double v) {
    __shared__ double s_part[32];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x * blockDim.y + 31) >> 5;
    v = warp_reduce<OP>(v);
    IMPEC_SYNCTHREADS();                 // protect s_part against a previous call in the same kernel
    if (lane == 0) s_part[warp] = v;
    IMPEC_SYNCTHREADS();
    if (warp == 0) {
        v = (lane < nwarps) ? s_part[lane] : red_identity<OP>();
        v = warp_reduce<OP>(v);
    }
    return v;
}
