// kernels_small.inl -- the bottom of the multigrid hierarchy in ONE launch.
//
// Levels of at most SM_MAXCELLS cells are latency-bound: a tile kernel there runs one or two CTAs for ~10 us
// (launch + a handful of dependent L2 round trips), and the W-cycle visits these levels eight times per cycle.
// k_mg_small walks a whole V-cycle over such levels (pre-smooth, residual, restriction ... coarsest solve ...
// prolongation, post-smooth) inside a single CTA of 1024 threads, with __syncthreads() where the tile kernels have
// kernel boundaries.  Iterates, right-hand sides and residuals of ALL these levels live in shared memory (~125 KB;
// iterates carry a zero border so the stencil needs no bounds checks); only the link coefficients are read from
// global memory (read-only, L1-resident after the first pass).  The arithmetic per cell is the one of k_mg_fused
// (same update order, same zero-guess identities, same transfer stencils), so the result differs from the tile
// path by rounding only.  Every colour pass is Jacobi inside the colour (new values staged in registers, barrier,
// store), exactly like the tile kernels.
#define SM_T 1024
#define SM_MAXLEV 8
#define SM_MAXCELLS 4096

struct SmallArgs { int nl, ncoarse; double *xout; LevelDev D[SM_MAXLEV]; };

// shared-memory footprint in doubles: bordered iterates + right-hand sides of all levels + one residual buffer
static inline size_t small_smem_doubles(const SmallArgs &A) {
    size_t n = (size_t)A.D[0].g.N1 * A.D[0].g.N2;
    for (int l = 0; l < A.nl; l++) n += (size_t)(A.D[l].g.N1 + 2) * (A.D[l].g.N2 + 2) + (size_t)A.D[l].g.N1 * A.D[l].g.N2;
    return n;
}

// sum_nb a_nb x_nb with x in the bordered shared layout (c = index of the cell, W = row pitch); k = global index of
// the cell in the coefficient arrays.  Order W,E,S,N,SW,SE,NW,NE as everywhere else.
template <class CF>
IMPEC_D double small_offsum(const CF &A, int ld, size_t k, const double *x, int c, int W) {
    double s = (double)A.AW[k] * x[c - W];
    s = fma((double)A.AW[k + ld], x[c + W], s);
    s = fma((double)A.AS[k], x[c - 1], s);
    s = fma((double)A.AS[k + 1], x[c + 1], s);
    s = fma((double)A.AV[k], x[c - W - 1], s);
    s = fma((double)A.AV[k + ld], x[c + W - 1], s);
    s = fma((double)A.AV[k + 1], x[c - W + 1], s);
    s = fma((double)A.AV[k + ld + 1], x[c + W + 1], s);
    return s;
}
template <class CF>
IMPEC_D double small_diagsum(const CF &A, int ld, size_t k, const double *x, int c, int W) {
    double s = (double)A.AV[k] * x[c - W - 1];
    s = fma((double)A.AV[k + ld], x[c + W - 1], s);
    s = fma((double)A.AV[k + 1], x[c - W + 1], s);
    s = fma((double)A.AV[k + ld + 1], x[c + W + 1], s);
    return s;
}

// Per-thread view of one level: the (at most SM_PER) cells this thread owns, with their indices computed once.
// NT = threads that walk the levels (1024 in k_mg_small, 512 inside the persistent tail kernel); tid = linear thread index
template <int NT> struct SmallCells {
    enum { PER = SM_MAXCELLS / NT };
    int c[PER];           // index in the bordered shared iterate (-1: no cell)
    int k[PER];           // index in the global coefficient arrays
    int t[PER];           // linear interior index (right-hand side / residual buffers)
    int par[PER];         // colour (i + j) & 1
};
template <int NT> IMPEC_D void small_cells(const Grid &g, SmallCells<NT> &Q, const int tid) {
    const int n = g.N1 * g.N2, W = g.N2 + 2;
    constexpr int SM_PER = SM_MAXCELLS / NT;
#pragma unroll
    for (int m = 0; m < SM_PER; m++) {
        const int t = tid + m * NT;
        if (t < n) {
            const int i = t / g.N2, j = t - i * g.N2;
            Q.c[m] = (i + 1) * W + j + 1; Q.k[m] = (int)gidx(g, i, j); Q.t[m] = t; Q.par[m] = (i + j) & 1;
        } else { Q.c[m] = -1; Q.k[m] = 0; Q.t[m] = 0; Q.par[m] = 0; }
    }
}

// one colour pass; ZERO: the iterate is zero on entry and is not read (first pass of a sweep from a zero guess)
template <bool F32, bool ZERO, int NT>
IMPEC_D void small_pass(const LevelDev &D, const SmallCells<NT> &Q, const double *f, double *x, const int col) {
    constexpr int SM_PER = SM_MAXCELLS / NT;
    const Coef<F32> A(D);
    const int ld = D.g.ld, W = D.g.N2 + 2;
    double nv[SM_PER];
#pragma unroll
    for (int m = 0; m < SM_PER; m++) {
        double v = 0.;
        if (Q.c[m] >= 0) {
            if (Q.par[m] == col) v = ZERO ? f[Q.t[m]] * (double)A.DINV[Q.k[m]] : (f[Q.t[m]] + small_offsum(A, ld, (size_t)Q.k[m], x, Q.c[m], W)) * (double)A.DINV[Q.k[m]];
            else if (!ZERO) v = x[Q.c[m]];
        }
        nv[m] = v;
    }
    if (!ZERO) IMPEC_SYNCTHREADS();
#pragma unroll
    for (int m = 0; m < SM_PER; m++)
        if (Q.c[m] >= 0) x[Q.c[m]] = nv[m];
    IMPEC_SYNCTHREADS();
}

// Coarsest level of at most 32 cells: one warp, one cell per lane, coefficients in registers, warp barriers only.
template <bool F32>
IMPEC_D void small_coarsest_warp(const LevelDev &D, const double *f, double *x, const int nsweeps, const int tid) {
    const Coef<F32> A(D);
    const Grid g = D.g;
    const int n = g.N1 * g.N2, W = g.N2 + 2, t = tid;
#ifdef IMPEC_EMU
    const bool in_warp = true;                    // the emulator's barriers are block-wide: every thread walks the loop
#else
    const bool in_warp = t < 32;
#endif
    if (in_warp) {
        const bool have = t < n;
        const int i = have ? t / g.N2 : 0, j = have ? t - i * g.N2 : 0, c = (i + 1) * W + j + 1, par = (i + j) & 1;
        const size_t k = gidx(g, i, j);
        const int ld = g.ld;
        double aw = 0., ae = 0., as = 0., an = 0., asw = 0., ase = 0., anw = 0., ane = 0., dinv = 0., fv = 0.;
        if (have) {
            aw = (double)A.AW[k]; ae = (double)A.AW[k + ld]; as = (double)A.AS[k]; an = (double)A.AS[k + 1];
            asw = (double)A.AV[k]; ase = (double)A.AV[k + ld]; anw = (double)A.AV[k + 1]; ane = (double)A.AV[k + ld + 1];
            dinv = (double)A.DINV[k]; fv = f[t];
        }
        for (int sw = 0; sw < nsweeps; sw++)
            for (int st = 0; st < 2; st++) {
                const int cfirst = sw & 1, col = st == 0 ? cfirst : 1 - cfirst;
                double v = 0.;
                const bool upd = have && par == col;
                if (upd) {
                    double s = aw * x[c - W];
                    s = fma(ae, x[c + W], s);
                    s = fma(as, x[c - 1], s);
                    s = fma(an, x[c + 1], s);
                    s = fma(asw, x[c - W - 1], s);
                    s = fma(ase, x[c + W - 1], s);
                    s = fma(anw, x[c - W + 1], s);
                    s = fma(ane, x[c + W + 1], s);
                    v = (fv + s) * dinv;
                }
                IMPEC_SYNCWARP();
                if (upd) x[c] = v;
                IMPEC_SYNCWARP();
            }
    }
    IMPEC_SYNCTHREADS();
}

// CG: the entry right-hand side was written by other CTAs of the same launch (persistent tail kernel): bypass L1
template <bool F32, int NT, bool CG>
IMPEC_D void small_body(const SmallArgs &A, const double *f0, double *S, const int tid) {
    constexpr int SM_PER = SM_MAXCELLS / NT;
    // carve shared memory: R | X_0 F_0 | X_1 F_1 | ...   (offsets in doubles, kept in shared memory)
    __shared__ int offX[SM_MAXLEV], offF[SM_MAXLEV];
    double *R = S;
    int total;
    {
        int p = A.D[0].g.N1 * A.D[0].g.N2;
        for (int l = 0; l < A.nl; l++) {
            if (tid == 0) { offX[l] = p; offF[l] = p + (A.D[l].g.N1 + 2) * (A.D[l].g.N2 + 2); }
            p += (A.D[l].g.N1 + 2) * (A.D[l].g.N2 + 2) + A.D[l].g.N1 * A.D[l].g.N2;
        }
        total = p;
    }
    // zero everything once (borders must be zero; iterates start from zero), then load the entry level's rhs
    for (int q = tid; q < total; q += NT) S[q] = 0.;
    IMPEC_SYNCTHREADS();
    {
        const Grid g = A.D[0].g;
        const int n = g.N1 * g.N2;
        double *F0 = S + offF[0];
        for (int t = tid; t < n; t += NT) { const int i = t / g.N2, j = t - i * g.N2; F0[t] = ldv<CG>(f0 + gidx(g, i, j)); }
    }
    IMPEC_SYNCTHREADS();
    SmallCells<NT> Q;
    // ---- down: forward sweep from zero (colour 0, then 1), residual, restriction
    for (int l = 0; l + 1 < A.nl; l++) {
        const LevelDev &D = A.D[l];
        const Grid g = D.g, gc = A.D[l + 1].g;
        const int W = g.N2 + 2;
        double *x = S + offX[l], *f = S + offF[l], *fc = S + offF[l + 1];
        small_cells<NT>(g, Q, tid);
        small_pass<F32, true, NT>(D, Q, f, x, 0);
        small_pass<F32, false, NT>(D, Q, f, x, 1);
        // residual after a sweep from zero: the first colour sees all eight neighbours, the second only its diagonals
        {
            const Coef<F32> C(D);
#pragma unroll
            for (int m = 0; m < SM_PER; m++) {
                if (Q.c[m] < 0) continue;
                double v = 0.;
                if (!is_pin(D, (long long)Q.k[m])) v = Q.par[m] == 0 ? small_offsum(C, g.ld, (size_t)Q.k[m], x, Q.c[m], W) : small_diagsum(C, g.ld, (size_t)Q.k[m], x, Q.c[m], W);
                R[Q.t[m]] = v;
            }
        }
        IMPEC_SYNCTHREADS();
        {
            const int nc = gc.N1 * gc.N2;
            for (int t = tid; t < nc; t += NT) {
                const int I = t / gc.N2, J = t - I * gc.N2;
                int jm = 2 * J - 1, jp = 2 * J + 2;
                jm = jm < 0 ? 0 : jm; jp = jp > g.N2 - 1 ? g.N2 - 1 : jp;
                double t4[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    int i = 2 * I - 1 + a;
                    i = i < 0 ? 0 : (i > g.N1 - 1 ? g.N1 - 1 : i);
                    const double *row = R + i * g.N2;
                    double s = 0.25 * row[jm];
                    s += 0.75 * row[2 * J];
                    s += 0.75 * row[2 * J + 1];
                    s += 0.25 * row[jp];
                    t4[a] = s;
                }
                fc[t] = 0.25 * (0.25 * t4[0] + 0.75 * t4[1] + 0.75 * t4[2] + 0.25 * t4[3]);
            }
        }
        IMPEC_SYNCTHREADS();
    }
    // ---- coarsest level: 2*ncoarse alternating sweeps from zero
    {
        const int l = A.nl - 1;
        const LevelDev &D = A.D[l];
        double *x = S + offX[l], *f = S + offF[l];
        if (D.g.N1 * D.g.N2 <= 32) small_coarsest_warp<F32>(D, f, x, 2 * A.ncoarse, tid);
        else {
            small_cells<NT>(D.g, Q, tid);
            small_pass<F32, true, NT>(D, Q, f, x, 0);
            small_pass<F32, false, NT>(D, Q, f, x, 1);
            for (int sw = 1; sw < 2 * A.ncoarse; sw++) {
                const int cfirst = sw & 1;
                small_pass<F32, false, NT>(D, Q, f, x, cfirst);
                small_pass<F32, false, NT>(D, Q, f, x, 1 - cfirst);
            }
        }
    }
    // ---- up: x += P e, backward sweep (colour 1, then 0)
    for (int l = A.nl - 2; l >= 0; l--) {
        const LevelDev &D = A.D[l];
        const Grid g = D.g, gc = A.D[l + 1].g;
        const int Wc = gc.N2 + 2;
        double *x = S + offX[l], *f = S + offF[l];
        const double *ec = S + offX[l + 1];
        small_cells<NT>(g, Q, tid);
#pragma unroll
        for (int m = 0; m < SM_PER; m++) {
            if (Q.c[m] < 0 || is_pin(D, (long long)Q.k[m])) continue;
            const int i = Q.t[m] / g.N2, j = Q.t[m] - i * g.N2;
            const int I = i >> 1, J = j >> 1;
            int In = (i & 1) ? I + 1 : I - 1, Jn = (j & 1) ? J + 1 : J - 1;
            In = In < 0 ? 0 : (In > gc.N1 - 1 ? gc.N1 - 1 : In);
            Jn = Jn < 0 ? 0 : (Jn > gc.N2 - 1 ? gc.N2 - 1 : Jn);
            const double *e0 = ec + (I + 1) * Wc + 1, *e1 = ec + (In + 1) * Wc + 1;
            const double pc = 0.75 * e0[J] + 0.25 * e1[J];
            const double pn = 0.75 * e0[Jn] + 0.25 * e1[Jn];
            x[Q.c[m]] += 0.75 * pc + 0.25 * pn;
        }
        IMPEC_SYNCTHREADS();
        small_pass<F32, false, NT>(D, Q, f, x, 1);
        small_pass<F32, false, NT>(D, Q, f, x, 0);
    }
    // ---- result of the entry level back to global memory (interior only; the padding of the array stays zero)
    {
        const Grid g = A.D[0].g;
        const int n = g.N1 * g.N2, W = g.N2 + 2;
        const double *x = S + offX[0];
        for (int t = tid; t < n; t += NT) { const int i = t / g.N2, j = t - i * g.N2; A.xout[gidx(g, i, j)] = x[(i + 1) * W + j + 1]; }
    }
}

template <bool F32>
__global__ void __launch_bounds__(SM_T) k_mg_small(SmallArgs A, const double *__restrict__ f0) {
    IMPEC_DYN_SMEM(double, S);
    small_body<F32, SM_T, false>(A, f0, S, (int)threadIdx.x);
}
