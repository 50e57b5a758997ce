// kernels_staged.inl -- second generation of the fused multigrid tile kernels.
//
// Same three fusions and the same arithmetic as k_mg_fused (kernels_fused.inl), but the coefficient tiles (FP32
// copies of AW, AS, AV, DINV) are staged into shared memory with cp.async (LDGSTS, L1 bypass) at kernel entry, so
// every global load of the CTA is in flight at once and the colour stages run entirely out of shared memory and
// registers.  k_mg_fused issues the coefficient loads row by row inside the stages and is bound by load latency
// (ncu: 15-18 warps stalled on long_scoreboard per issue, profiles/r1e_*); this version is meant to be bound by HBM.
// Tiles are 32 x 128 cells (halo included) so that x (32 KB) plus four coefficient tiles (67 KB) allow two CTAs per SM.
#define ST_H 32                          // tile rows, halo included
#define ST_NTY 8                         // thread rows per CTA
#define ST_ROWS (ST_H / ST_NTY)          // rows per thread (4)
#define ST_CP 132                        // pitch of the coefficient tiles in floats (128 + the column to the right, 16 B chunks)
template <int MODE> struct StagedGeom { enum { HI = (MODE == FUSED_PRE) ? 3 : 2, TI = ST_H - 2 * HI }; };
#define ST_SMEM_BYTES (ST_H * FT_W * 8 + ((ST_H + 1) + ST_H + (ST_H + 1) + ST_H) * ST_CP * 4)

struct StagedTiles { double *X; float *AW, *AS, *AV, *DI; };
IMPEC_D StagedTiles staged_carve(unsigned char *raw) {
    StagedTiles t;
    t.X = reinterpret_cast<double *>(raw);
    t.AW = reinterpret_cast<float *>(raw + ST_H * FT_W * 8);
    t.AS = t.AW + (ST_H + 1) * ST_CP;
    t.AV = t.AS + ST_H * ST_CP;
    t.DI = t.AV + (ST_H + 1) * ST_CP;
    return t;
}

// sum_nb a_nb x_nb for tile cell (r, jj = 2 tx + off), coefficients from the staged tiles
IMPEC_D double staged_offsum(const StagedTiles &T, int r, int tx, int off) {
    const int jj = 2 * tx + off;
    const double *own = T.X + r * FT_W + (off << 6) + tx;
    const double *oth = T.X + r * FT_W + ((1 - off) << 6) + tx + off;
    const float *aw = T.AW + r * ST_CP + jj, *as = T.AS + r * ST_CP + jj, *av = T.AV + r * ST_CP + jj;
    double s = (double)aw[0] * own[-FT_W];
    s = fma((double)aw[ST_CP], own[FT_W], s);
    s = fma((double)as[0], oth[-1], s);
    s = fma((double)as[1], oth[0], s);
    s = fma((double)av[0], oth[-FT_W - 1], s);
    s = fma((double)av[ST_CP], oth[FT_W - 1], s);
    s = fma((double)av[1], oth[-FT_W], s);
    s = fma((double)av[ST_CP + 1], oth[FT_W], s);
    return s;
}
IMPEC_D double staged_diagsum(const StagedTiles &T, int r, int tx, int off) {
    const int jj = 2 * tx + off;
    const double *oth = T.X + r * FT_W + ((1 - off) << 6) + tx + off;
    const float *av = T.AV + r * ST_CP + jj;
    double s = (double)av[0] * oth[-FT_W - 1];
    s = fma((double)av[ST_CP], oth[FT_W - 1], s);
    s = fma((double)av[1], oth[-FT_W], s);
    s = fma((double)av[ST_CP + 1], oth[FT_W], s);
    return s;
}

// asynchronous copy of `nrows` x ST_CP floats of a coefficient array into its tile; chunks outside the allocation are zeroed
IMPEC_D void staged_fetch(float *tile, const float *__restrict__ src, const Grid &g, int i0, int j0, int nrows, int tid, int nthreads) {
    const int chunks = ST_CP / 4;                                            // 16-byte chunks per row
    for (int t = tid; t < nrows * chunks; t += nthreads) {
        const int r = t / chunks, cch = t - r * chunks;
        const int ig = i0 + r, col = j0 + IMPEC_PJ + 4 * cch;               // first float of the chunk within the padded row
        float *dst = tile + r * ST_CP + 4 * cch;
        if (ig >= -IMPEC_GH && ig < g.rows - IMPEC_GH && col >= 0 && col + 3 < g.ld)
            impec_cp_async16(dst, src + (size_t)(ig + IMPEC_GH) * g.ld + col);
        else { dst[0] = 0.f; dst[1] = 0.f; dst[2] = 0.f; dst[3] = 0.f; }
    }
}

template <int MODE, bool EDGE>
IMPEC_D void staged_body(const LevelDev &L, const double *__restrict__ f, const double *__restrict__ xin, double *__restrict__ x, const int c0,
                         const Grid &gc, const double *__restrict__ ec, double *__restrict__ fc, double *__restrict__ part, const StagedTiles &T) {
    constexpr int HI = StagedGeom<MODE>::HI, TI = StagedGeom<MODE>::TI;
    const Grid g = L.g;
    const int ld = g.ld;
    double *X = T.X;
    const int tx = threadIdx.x, ty = threadIdx.y & (ST_NTY - 1);
    const int tid = ty * (FT_W / 2) + tx, nthreads = (FT_W / 2) * ST_NTY;
    const int jl = 2 * tx;
    const int j0 = (int)blockIdx.x * FT_TJ - FT_HJ;
    const int i0 = (int)blockIdx.y * TI - HI;
    const int jg = j0 + jl;
    const int par = (i0 + ty + jg) & 1;
    const long long kbase = (long long)(i0 + ty + IMPEC_GH) * ld + (jg + IMPEC_PJ);
    const bool col_alloc = !EDGE || ((jg + IMPEC_PJ >= 0) && (jg + 1 + IMPEC_PJ < ld));

    // ---- everything this CTA reads from global memory is requested here, before any of it is used
    staged_fetch(T.AW, L.AWf, g, i0, j0, ST_H + 1, tid, nthreads);
    staged_fetch(T.AS, L.ASf, g, i0, j0, ST_H, tid, nthreads);
    staged_fetch(T.AV, L.AVf, g, i0, j0, ST_H + 1, tid, nthreads);
    staged_fetch(T.DI, L.DINVf, g, i0, j0, ST_H, tid, nthreads);
    double fa[ST_ROWS], fb[ST_ROWS];                                          // right-hand side of this thread's pairs
#pragma unroll
    for (int n = 0; n < ST_ROWS; n++) {
        const int r = ty + n * ST_NTY, ig = i0 + r;
        double a = 0., b = 0.;
        fa[n] = 0.; fb[n] = 0.;
        const bool row_alloc = !EDGE || ((ig >= -IMPEC_GH) && (ig < g.rows - IMPEC_GH));
        if (row_alloc && col_alloc) {
            const size_t k = (size_t)(kbase + (long long)n * ST_NTY * ld);
            const d2 fv = ld2(f + k); fa[n] = fv.x; fb[n] = fv.y;
            if (MODE != FUSED_PRE && xin) {
                const d2 v = ld2(xin + k);
                a = v.x; b = v.y;
                if (MODE == FUSED_POST && (!EDGE || (ig >= g.rlo && ig < g.rhi))) {
                    const int I = ig >> 1, J = jg >> 1;
                    int In = (ig & 1) ? I + 1 : I - 1, Jm = J - 1, Jp = J + 1;
                    if (EDGE) {
                        if (In < 0 && gc.wall_lo) In = 0;
                        if (In > gc.N1 - 1 && gc.wall_hi) In = gc.N1 - 1;
                        Jm = Jm < 0 ? 0 : Jm; Jp = Jp > gc.N2 - 1 ? gc.N2 - 1 : Jp;
                    }
                    const double *e0 = ec + gidx(gc, I, 0), *e1 = ec + gidx(gc, In, 0);
                    const double pc = 0.75 * e0[J] + 0.25 * e1[J];
                    const double pm = 0.75 * e0[Jm] + 0.25 * e1[Jm];
                    const double pp = 0.75 * e0[Jp] + 0.25 * e1[Jp];
                    if (!EDGE) { a += 0.75 * pc + 0.25 * pm; b += 0.75 * pc + 0.25 * pp; }
                    else {
                        if (jg >= 0 && jg < g.N2 && (long long)k != L.pin) a += 0.75 * pc + 0.25 * pm;
                        if (jg + 1 >= 0 && jg + 1 < g.N2 && (long long)(k + 1) != L.pin) b += 0.75 * pc + 0.25 * pp;
                    }
                }
            }
        }
        X[r * FT_W + tx] = a; X[r * FT_W + 64 + tx] = b;
    }
    impec_cp_async_wait_all();
    IMPEC_SYNCTHREADS();

    // ---- two colour stages out of shared memory
    double nv[ST_ROWS];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int col = s == 0 ? c0 : 1 - c0;
        const int shrink = (MODE == FUSED_PRE) ? s : s + 1;
        const int off = (par == col) ? 0 : 1;
        const int jj = jl + off, j = jg + off;
        const bool colok = jj >= shrink && jj < FT_W - shrink && (!EDGE || (j >= 0 && j < g.N2));
#pragma unroll
        for (int n = 0; n < ST_ROWS; n++) {
            const int r = ty + n * ST_NTY, ig = i0 + r;
            double val = X[r * FT_W + (off << 6) + tx];
            const bool act = colok && r >= shrink && r < ST_H - shrink && (!EDGE || (ig >= g.rlo && ig < g.rhi));
            if (act) {
                const double fv = off ? fb[n] : fa[n];
                const double di = (double)T.DI[r * ST_CP + jj];
                if (MODE == FUSED_PRE && s == 0) val = fv * di;
                else val = (fv + staged_offsum(T, r, tx, off)) * di;
            }
            nv[n] = val;
        }
        if (!(MODE == FUSED_PRE && s == 0)) IMPEC_SYNCTHREADS();
#pragma unroll
        for (int n = 0; n < ST_ROWS; n++) X[(ty + n * ST_NTY) * FT_W + (off << 6) + tx] = nv[n];
        IMPEC_SYNCTHREADS();
    }

    // ---- interior write-back and the fused dot product
    const bool col_int = jl >= FT_HJ && jl < FT_W - FT_HJ && (!EDGE || jg < g.N2);
    double dotv = 0.;
#pragma unroll
    for (int n = 0; n < ST_ROWS; n++) {
        const int r = ty + n * ST_NTY, ig = i0 + r;
        if (r < HI || r >= ST_H - HI || !col_int || (EDGE && ig >= g.N1)) continue;
        const size_t k = (size_t)(kbase + (long long)n * ST_NTY * ld);
        const double a = X[r * FT_W + tx], b = X[r * FT_W + 64 + tx];
        if (!EDGE || jg + 1 < g.N2) st2(x + k, a, b); else x[k] = a;
        if (MODE == FUSED_POST && part) { dotv = fma(fa[n], a, dotv); if (!EDGE || jg + 1 < g.N2) dotv = fma(fb[n], b, dotv); }
    }
    if (MODE == FUSED_POST && part) {
        dotv = block_reduce<RED_SUM>(dotv);
        if (tx == 0 && ty == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = dotv;
    }
    if (MODE != FUSED_PRE) return;

    // ---- PRE: residual (zero-guess identity, see kernels_fused.inl) and full-weighting restriction
    double ra[ST_ROWS], rb[ST_ROWS];
    {
        const int off0 = (par == c0) ? 0 : 1;
#pragma unroll
        for (int n = 0; n < ST_ROWS; n++) {
            const int r = ty + n * ST_NTY, ig = i0 + r;
            double v0 = 0., v1 = 0.;
            if (r >= 2 && r < ST_H - 2 && (!EDGE || (ig >= g.rlo && ig < g.rhi))) {
                const size_t k = (size_t)(kbase + (long long)n * ST_NTY * ld);
                const int ja = jl + off0, jb = jl + 1 - off0;
                if (ja >= 2 && ja < FT_W - 2 && (!EDGE || (jg + off0 >= 0 && jg + off0 < g.N2 && (long long)(k + off0) != L.pin)))
                    v0 = staged_offsum(T, r, tx, off0);
                if (jb >= 2 && jb < FT_W - 2 && (!EDGE || (jg + 1 - off0 >= 0 && jg + 1 - off0 < g.N2 && (long long)(k + 1 - off0) != L.pin)))
                    v1 = staged_diagsum(T, r, tx, 1 - off0);
            }
            ra[n] = off0 == 0 ? v0 : v1;
            rb[n] = off0 == 0 ? v1 : v0;
        }
    }
    IMPEC_SYNCTHREADS();
#pragma unroll
    for (int n = 0; n < ST_ROWS; n++) {
        const int r = ty + n * ST_NTY;
        X[r * FT_W + tx] = ra[n]; X[r * FT_W + 64 + tx] = rb[n];
    }
    IMPEC_SYNCTHREADS();
    const int I0 = (i0 + HI) >> 1, J0 = (j0 + FT_HJ) >> 1;
    for (int t = tid; t < (TI / 2) * (FT_TJ / 2); t += nthreads) {
        const int I = I0 + t / (FT_TJ / 2), J = J0 + t % (FT_TJ / 2);
        if (EDGE && (I >= gc.N1 || J >= gc.N2)) continue;
        double t4[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            int i = 2 * I - 1 + a;
            if (EDGE) { if (i < 0 && g.wall_lo) i = 0; if (i > g.N1 - 1 && g.wall_hi) i = g.N1 - 1; }
            int jm = 2 * J - 1, jp = 2 * J + 2;
            if (EDGE) { jm = jm < 0 ? 0 : jm; jp = jp > g.N2 - 1 ? g.N2 - 1 : jp; }
            const int rr = i - i0;
            double s = 0.25 * X[xi(rr, jm - j0)];
            s += 0.75 * X[xi(rr, 2 * J - j0)];
            s += 0.75 * X[xi(rr, 2 * J + 1 - j0)];
            s += 0.25 * X[xi(rr, jp - j0)];
            t4[a] = s;
        }
        fc[gidx(gc, I, J)] = 0.25 * (0.25 * t4[0] + 0.75 * t4[1] + 0.75 * t4[2] + 0.25 * t4[3]);
    }
}

template <int MODE>
__global__ void __launch_bounds__(FT_W / 2 * ST_NTY, 2) k_mg_staged(LevelDev L, const double *__restrict__ f, const double *__restrict__ xin, double *__restrict__ x, int c0,
                                                                   Grid gc, const double *__restrict__ ec, double *__restrict__ fc, double *__restrict__ part) {
    IMPEC_DYN_SMEM(unsigned char, raw);
    const StagedTiles T = staged_carve(raw);
    constexpr int HI = StagedGeom<MODE>::HI, TI = StagedGeom<MODE>::TI;
    const int j0 = (int)blockIdx.x * FT_TJ - FT_HJ, i0 = (int)blockIdx.y * TI - HI;
    bool edge = i0 < L.g.rlo || j0 < 0 || i0 + ST_H >= L.g.rhi || j0 + FT_W >= L.g.N2;
    if (L.pin >= 0) {
        const int pi = (int)(L.pin / L.g.ld) - IMPEC_GH, pj = (int)(L.pin % L.g.ld) - IMPEC_PJ;
        edge = edge || (pi >= i0 && pi < i0 + ST_H && pj >= j0 && pj < j0 + FT_W);
    }
    if (edge) staged_body<MODE, true>(L, f, xin, x, c0, gc, ec, fc, part, T);
    else staged_body<MODE, false>(L, f, xin, x, c0, gc, ec, fc, part, T);
}
