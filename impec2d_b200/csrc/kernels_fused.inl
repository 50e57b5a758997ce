// kernels_fused.inl -- fused multigrid tile kernels (the hot kernels of the pressure solve).
//
// One CTA stages a 64 x 128 tile of the iterate (halo included) in shared memory and performs BOTH colour
// passes of the red-black sweep on it, so a sweep costs one read and one write of x instead of two each.
// Halo cells are recomputed redundantly (identical arithmetic, so tiles agree bit for bit).
//
//   FUSED_PLAIN : x <- sweep(x)                                   (coarsest level, W-cycle mid sweeps)
//   FUSED_PRE   : x <- sweep(0); r = f - S x; fc = R r            (pre-smooth from a zero guess + residual +
//                                                                  full-weighting restriction; x is never read)
//   FUSED_POST  : x <- x + P ec; x <- sweep(x); part += <f,x>     (prolongation + correction + post-smooth,
//                                                                  with the PCG dot product <r,z> on level 0)
//   FUSED_MID   : x <- sweep(x); r = f - S x; fc = R r            (second visit of a W-cycle level: sweep from the current
//                                                                  iterate + full residual + restriction; single-GPU levels)
//
// Thread (tx,ty) owns the cell pair (columns 2tx, 2tx+1) of rows ty, ty+8, ...: one cell of the pair is red
// and one black, so both colour stages keep every lane busy and global accesses are 128-bit.
// Shared memory holds x only, split into an even-column and an odd-column plane so that every access of a
// colour stage is unit-stride across the warp (no bank conflicts).  Link coefficients and the stored inverse
// diagonal are read through L1.  Tiles that touch the domain boundary (or hold the pinned cell) take the EDGE
// instantiation with full bounds checks; all others run the check-free body.
#define FT_W 128                 // tile width in cells, halo included
#define FT_H 64                  // tile height in cells, halo included
#define FT_HJ 4                  // halo in j (keeps 32-byte alignment of the tile's first column)
#define FT_NTY 8                 // thread rows per CTA
#define FT_ROWS (FT_H / FT_NTY)  // rows per thread
#define FT_TJ (FT_W - 2 * FT_HJ) // interior columns per tile (120)
enum { FUSED_PLAIN = 0, FUSED_PRE = 1, FUSED_POST = 2, FUSED_MID = 3 };
// TH = tile height: 64 on the big levels (least halo overhead); 32 / 16 on the small ones, where a level is only a few
// tiles and the time of a launch is the time ONE CTA needs for its tile (more, shorter CTAs finish sooner)
template <int MODE, int TH = FT_H> struct FusedGeom { enum { HI = (MODE == FUSED_PRE) ? 3 : ((MODE == FUSED_MID) ? 4 : 2), TI = TH - 2 * HI }; };

// loads of ITERATES / right-hand sides.  CG = true bypasses L1 (ld.global.cg): needed inside the persistent tail kernel
// (k_mg_tail), where one CTA reads arrays that other CTAs rewrite between the phases of the SAME launch -- L1 is only
// coherent at launch boundaries.  Link coefficients never change during a launch and always use plain loads.
template <bool CG> IMPEC_D double ldv(const double *p) {
#ifndef IMPEC_EMU
    return CG ? __ldcg(p) : *p;
#else
    return *p;
#endif
}
template <bool CG> IMPEC_D d2 ldv2(const double *p) {
#ifndef IMPEC_EMU
    if (CG) { const double2 v = __ldcg(reinterpret_cast<const double2 *>(p)); d2 r; r.x = v.x; r.y = v.y; return r; }
#endif
    return ld2(p);
}

// shared-memory index of tile cell (r, jj): row-major, even columns in the first half of the row
IMPEC_D int xi(int r, int jj) { return r * FT_W + ((jj & 1) << 6) + (jj >> 1); }

// coefficient views: F32 = the cycle's FP32 copies, otherwise the FP64 arrays
template <bool F32> struct Coef;
template <> struct Coef<false> {
    const double *AW, *AS, *AV, *DINV;
    IMPEC_D Coef(const LevelDev &L) : AW(L.AW), AS(L.AS), AV(L.AV), DINV(L.DINV) {}
};
template <> struct Coef<true> {
    const float *AW, *AS, *AV, *DINV;
    IMPEC_D Coef(const LevelDev &L) : AW(L.AWf), AS(L.ASf), AV(L.AVf), DINV(L.DINVf) {}
};

// sum_nb a_nb x_nb for the cell `off` of the pair owned by lane tx in tile row r (k = its global index)
template <class CF>
IMPEC_D double tile_offsum(const CF &A, int ld, const double *__restrict__ X, int r, int tx, int off, size_t k) {
    const double *own = X + r * FT_W + (off << 6) + tx;                  // same column: rows r-1, r+1
    const double *oth = X + r * FT_W + ((1 - off) << 6) + tx + off;      // other plane: [-1] is column jj-1, [0] is jj+1
    double s = (double)A.AW[k] * own[-FT_W];
    s = fma((double)A.AW[k + ld], own[FT_W], s);
    s = fma((double)A.AS[k], oth[-1], s);
    s = fma((double)A.AS[k + 1], oth[0], s);
    s = fma((double)A.AV[k], oth[-FT_W - 1], s);
    s = fma((double)A.AV[k + ld], oth[FT_W - 1], s);
    s = fma((double)A.AV[k + 1], oth[-FT_W], s);
    s = fma((double)A.AV[k + ld + 1], oth[FT_W], s);
    return s;
}
// the four diagonal terms only
template <class CF>
IMPEC_D double tile_diagsum(const CF &A, int ld, const double *__restrict__ X, int r, int tx, int off, size_t k) {
    const double *oth = X + r * FT_W + ((1 - off) << 6) + tx + off;
    double s = (double)A.AV[k] * oth[-FT_W - 1];
    s = fma((double)A.AV[k + ld], oth[FT_W - 1], s);
    s = fma((double)A.AV[k + 1], oth[-FT_W], s);
    s = fma((double)A.AV[k + ld + 1], oth[FT_W], s);
    return s;
}

// diagonal of the links operator from the coefficient view (same association as link_diag)
template <class CF>
IMPEC_D double tile_diag(const CF &A, const LevelDev &L, int ld, size_t k) {
    double d = L.mass + ((((double)A.AW[k] + (double)A.AW[k + ld]) + ((double)A.AS[k] + (double)A.AS[k + 1])) +
                         (((double)A.AV[k] + (double)A.AV[k + ld]) + ((double)A.AV[k + 1] + (double)A.AV[k + ld + 1])));
    if (has_sig(L, (long long)k)) d += L.sig[(long long)k - L.sig_idx];
    return d;
}

// nsweeps > 1 (PLAIN only, and only when this single tile holds the WHOLE level so that no halo can go stale): the
// sweeps alternate forward/backward starting with colour c0 -- the coarsest-level solve in one launch.
// (bx, by): tile coordinates, nbx: tiles per row of tiles (the tile kernels pass blockIdx / gridDim; the persistent tail kernel
// loops over tiles).  The restrict qualifiers are dropped for CG: those arrays are written by other CTAs during the launch.
template <int MODE, bool EDGE, bool F32, int TH, bool CG = false>
IMPEC_D void fused_body(const LevelDev &L, const double *f, const double *xin, double *x, const int c0,
                        const Grid &gc, const double *ec, double *fc, double *part, double *__restrict__ X, const int nsweeps,
                        const int bx, const int by, const int nbx) {
    constexpr int HI = FusedGeom<MODE, TH>::HI, TI = FusedGeom<MODE, TH>::TI;
    const Grid g = L.g;
    const int ld = g.ld;
    const Coef<F32> A(L);
    const int tx = threadIdx.x, ty = threadIdx.y & (FT_NTY - 1);
    const int jl = 2 * tx;                                   // local column of the pair
    const int j0 = bx * FT_TJ - FT_HJ;                       // global column of tile column 0 (even)
    const int i0 = by * TI - HI;                             // global row of tile row 0
    const int jg = j0 + jl;
    const int par = (i0 + ty + jg) & 1;                      // colour of the pair's first cell (same in all rows of this thread)
    const long long kbase = (long long)(i0 + ty + IMPEC_GH) * ld + (jg + IMPEC_PJ);   // global index of the pair in row ty
    const bool col_alloc = !EDGE || ((jg + IMPEC_PJ >= 0) && (jg + 1 + IMPEC_PJ < ld));

    // ---- stage L: tile of x (zero for a zero initial guess), plus the prolongated correction
#pragma unroll
    for (int n = 0; n < (TH / FT_NTY); n++) {
        const int r = ty + n * FT_NTY, ig = i0 + r;
        double a = 0., b = 0.;
        const bool row_alloc = !EDGE || ((ig >= -IMPEC_GH) && (ig < g.rows - IMPEC_GH));
        if (MODE != FUSED_PRE && row_alloc && col_alloc && xin) {
            const size_t k = (size_t)(kbase + (long long)n * FT_NTY * ld);
            const d2 v = ldv2<CG>(xin + k);
            a = v.x; b = v.y;
            if (MODE == FUSED_POST && (!EDGE || (ig >= g.rlo && ig < g.rhi))) {
                // x += P ec : cell-centred bilinear with edge replication.  Both cells of the pair share the coarse
                // column J; the even cell interpolates towards J-1, the odd one towards J+1.
                const int I = ig >> 1, J = jg >> 1;
                int In = (ig & 1) ? I + 1 : I - 1, Jm = J - 1, Jp = J + 1;
                if (EDGE) {
                    if (In < 0 && gc.wall_lo) In = 0;
                    if (In > gc.N1 - 1 && gc.wall_hi) In = gc.N1 - 1;
                    Jm = Jm < 0 ? 0 : Jm; Jp = Jp > gc.N2 - 1 ? gc.N2 - 1 : Jp;
                }
                const double *e0 = ec + gidx(gc, I, 0), *e1 = ec + gidx(gc, In, 0);
                const double pc = 0.75 * ldv<CG>(e0 + J) + 0.25 * ldv<CG>(e1 + J);
                const double pm = 0.75 * ldv<CG>(e0 + Jm) + 0.25 * ldv<CG>(e1 + Jm);
                const double pp = 0.75 * ldv<CG>(e0 + Jp) + 0.25 * ldv<CG>(e1 + Jp);
                if (!EDGE) { a += 0.75 * pc + 0.25 * pm; b += 0.75 * pc + 0.25 * pp; }
                else {
                    if (jg >= 0 && jg < g.N2 && !is_pin(L, (long long)k)) a += 0.75 * pc + 0.25 * pm;
                    if (jg + 1 >= 0 && jg + 1 < g.N2 && !is_pin(L, (long long)(k + 1))) b += 0.75 * pc + 0.25 * pp;
                }
            }
        }
        X[r * FT_W + tx] = a; X[r * FT_W + 64 + tx] = b;
    }
    IMPEC_SYNCTHREADS();
    // (xin and x are different buffers for PLAIN / POST: the halo of this tile is the interior of its
    //  neighbours, which other CTAs overwrite concurrently.)

    // ---- two colour stages.  Stage s updates colour (s == 0 ? c0 : 1 - c0) on the tile shrunk by s+1 cells
    //      (by s cells in PRE mode, whose first stage is the neighbour-free zero-guess update x = f/diag).
    //      The pinned cell has dinv = 0 and x = 0, so it needs no special case here.
    double nv[(TH / FT_NTY)];
    for (int sw = 0; sw < nsweeps; sw++)
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int cfirst = (sw & 1) ? 1 - c0 : c0;
        const int col = s == 0 ? cfirst : 1 - cfirst;
        const int shrink = (MODE == FUSED_PRE) ? s : s + 1;
        const int off = (par == col) ? 0 : 1;                            // which cell of the pair has this colour
        const int jj = jl + off, j = jg + off;
        const bool colok = jj >= shrink && jj < FT_W - shrink && (!EDGE || (j >= 0 && j < g.N2));
#pragma unroll
        for (int n = 0; n < (TH / FT_NTY); n++) {
            const int r = ty + n * FT_NTY, ig = i0 + r;
            double val = X[r * FT_W + (off << 6) + tx];
            const bool act = colok && r >= shrink && r < TH - shrink && (!EDGE || (ig >= g.rlo && ig < g.rhi));
            if (act) {
                const size_t k = (size_t)(kbase + (long long)n * FT_NTY * ld + off);
                if (MODE == FUSED_PRE && s == 0) val = ldv<CG>(f + k) * (double)A.DINV[k];
                else val = (ldv<CG>(f + k) + tile_offsum(A, ld, X, r, tx, off, k)) * (double)A.DINV[k];
            }
            nv[n] = val;
        }
        if (!(MODE == FUSED_PRE && s == 0)) IMPEC_SYNCTHREADS();        // every read of the old colour is done
#pragma unroll
        for (int n = 0; n < (TH / FT_NTY); n++) X[(ty + n * FT_NTY) * FT_W + (off << 6) + tx] = nv[n];
        IMPEC_SYNCTHREADS();
    }

    // ---- write the interior of the tile back (128-bit stores), and the fused dot product
    const bool col_int = jl >= FT_HJ && jl < FT_W - FT_HJ && (!EDGE || jg < g.N2);
    double dotv = 0.;
    // PRE on a slab: f is valid on 3 halo rows of the neighbouring slabs, so the sweep from zero is exact on 2 halo
    // rows as well.  The first / last row of tiles writes them too, which saves the exchange of x that the
    // post-smoothing kernel would otherwise need.
    const int rmin_w = (MODE == FUSED_PRE && by == 0 && !g.wall_lo) ? HI - 2 : HI;
#pragma unroll
    for (int n = 0; n < (TH / FT_NTY); n++) {
        const int r = ty + n * FT_NTY, ig = i0 + r;
        const bool top_halo = MODE == FUSED_PRE && EDGE && !g.wall_hi && ig >= g.N1 && ig < g.N1 + 2 && r < TH - 1;
        if (r < rmin_w || !col_int) continue;
        if (!top_halo && (r >= TH - HI || (EDGE && ig >= g.N1))) continue;
        const size_t k = (size_t)(kbase + (long long)n * FT_NTY * ld);
        const double a = X[r * FT_W + tx], b = X[r * FT_W + 64 + tx];
        if (!EDGE || jg + 1 < g.N2) st2(x + k, a, b); else x[k] = a;
        if (MODE == FUSED_POST && part) {
            if (!EDGE || jg + 1 < g.N2) { const d2 fv = ldv2<CG>(f + k); dotv = fma(fv.x, a, dotv); dotv = fma(fv.y, b, dotv); }
            else dotv = fma(ldv<CG>(f + k), a, dotv);
        }
    }
    if (MODE == FUSED_POST && part) {
        dotv = block_reduce<RED_SUM>(dotv);
        if (tx == 0 && ty == 0) part[by * nbx + bx] = dotv;
    }
    if (MODE != FUSED_PRE && MODE != FUSED_MID) return;

    // ---- PRE: residual on the tile shrunk by 2, then full-weighting restriction.
    // After a sweep from a ZERO guess the residual needs no diagonal: a first-colour cell was solved when all
    // its neighbours were still zero, so r = sum over all 8 neighbours of a_nb x_nb; a second-colour cell was
    // solved against the final first-colour values, so only its (same-colour) diagonal neighbours contribute.
    // (The dropped terms f - d x + s_at_update are rounding noise, O(eps |f|).)
    // ---- MID: the sweep started from a non-zero iterate, so the residual is evaluated in full, on the tile shrunk by 3.
    double ra[(TH / FT_NTY)], rb[(TH / FT_NTY)];
    if (MODE == FUSED_MID) {
#pragma unroll
        for (int n = 0; n < (TH / FT_NTY); n++) {
            const int r = ty + n * FT_NTY, ig = i0 + r;
            double v[2] = {0., 0.};
            if (r >= 3 && r < TH - 3 && (!EDGE || (ig >= g.rlo && ig < g.rhi))) {
                const size_t k = (size_t)(kbase + (long long)n * FT_NTY * ld);
#pragma unroll
                for (int off = 0; off < 2; off++) {
                    const int jc = jl + off;
                    if (jc >= 3 && jc < FT_W - 3 && (!EDGE || (jg + off >= 0 && jg + off < g.N2 && !is_pin(L, (long long)(k + off)))))
                        v[off] = ldv<CG>(f + k + off) - (tile_diag(A, L, ld, k + off) * X[r * FT_W + (off << 6) + tx] - tile_offsum(A, ld, X, r, tx, off, k + off));
                }
            }
            ra[n] = v[0]; rb[n] = v[1];
        }
    } else {
        const int off0 = (par == c0) ? 0 : 1;                            // first-colour cell of the pair
#pragma unroll
        for (int n = 0; n < (TH / FT_NTY); n++) {
            const int r = ty + n * FT_NTY, ig = i0 + r;
            double v0 = 0., v1 = 0.;                                     // residual of the first / second colour cell
            if (r >= 2 && r < TH - 2 && (!EDGE || (ig >= g.rlo && ig < g.rhi))) {
                const size_t k = (size_t)(kbase + (long long)n * FT_NTY * ld);
                const int ja = jl + off0, jb = jl + 1 - off0;
                if (ja >= 2 && ja < FT_W - 2 && (!EDGE || (jg + off0 >= 0 && jg + off0 < g.N2 && !is_pin(L, (long long)(k + off0)))))
                    v0 = tile_offsum(A, ld, X, r, tx, off0, k + off0);
                if (jb >= 2 && jb < FT_W - 2 && (!EDGE || (jg + 1 - off0 >= 0 && jg + 1 - off0 < g.N2 && !is_pin(L, (long long)(k + 1 - off0)))))
                    v1 = tile_diagsum(A, ld, X, r, tx, 1 - off0, k + 1 - off0);
            }
            ra[n] = off0 == 0 ? v0 : v1;                                 // even-column cell
            rb[n] = off0 == 0 ? v1 : v0;                                 // odd-column cell
        }
    }
    IMPEC_SYNCTHREADS();
#pragma unroll
    for (int n = 0; n < (TH / FT_NTY); n++) {
        const int r = ty + n * FT_NTY;
        X[r * FT_W + tx] = ra[n]; X[r * FT_W + 64 + tx] = rb[n];
    }
    IMPEC_SYNCTHREADS();
    // coarse cells of this tile: fine interior rows [i0+HI, i0+HI+TI) -> TI/2 coarse rows, FT_TJ/2 coarse columns
    const int I0 = (i0 + HI) >> 1, J0 = (j0 + FT_HJ) >> 1;
    const int tid = ty * (FT_W / 2) + tx, nthreads = (FT_W / 2) * FT_NTY;
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
            // 0.25 r(2J-1) + 0.75 r(2J) + 0.75 r(2J+1) + 0.25 r(2J+2)
            double s = 0.25 * X[xi(rr, jm - j0)];
            s += 0.75 * X[xi(rr, 2 * J - j0)];
            s += 0.75 * X[xi(rr, 2 * J + 1 - j0)];
            s += 0.25 * X[xi(rr, jp - j0)];
            t4[a] = s;
        }
        fc[gidx(gc, I, J)] = 0.25 * (0.25 * t4[0] + 0.75 * t4[1] + 0.75 * t4[2] + 0.25 * t4[3]);
    }
}

template <int MODE, bool F32, int TH>
__global__ void __launch_bounds__(FT_W / 2 * FT_NTY, (TH == 64 ? 2 : (TH == 32 ? 3 : 4))) k_mg_fused(LevelDev L, const double *__restrict__ f, const double *__restrict__ xin, double *__restrict__ x, int c0,
                                                               Grid gc, const double *__restrict__ ec, double *__restrict__ fc, double *__restrict__ part) {
    IMPEC_DYN_SMEM(double, X);
    constexpr int HI = FusedGeom<MODE, TH>::HI, TI = FusedGeom<MODE, TH>::TI;
    const int j0 = (int)blockIdx.x * FT_TJ - FT_HJ, i0 = (int)blockIdx.y * TI - HI;
    // tile (halo included, plus the one-cell reach of the link arrays) strictly inside the domain and free of the pin?
    bool edge = i0 < L.g.rlo || j0 < 0 || i0 + TH >= L.g.rhi || j0 + FT_W >= L.g.N2;
    if (L.pin >= 0) {                                   // the pinned cells are one cell or (part of) one row
        const int pi = (int)(L.pin / L.g.ld) - IMPEC_GH, pj = (int)(L.pin % L.g.ld) - IMPEC_PJ, pje = pj + (int)(L.pin_end - L.pin);
        edge = edge || (pi >= i0 && pi < i0 + TH && pje > j0 && pj < j0 + FT_W);
    }
    if (edge) fused_body<MODE, true, F32, TH>(L, f, xin, x, c0, gc, ec, fc, part, X, 1, (int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x);
    else fused_body<MODE, false, F32, TH>(L, f, xin, x, c0, gc, ec, fc, part, X, 1, (int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x);
}

// Coarsest-level solve: `nsweeps` alternating red-black sweeps from a zero guess in ONE launch.  Requires that a
// single tile holds the whole level (N1 <= FT_H - 4, N2 <= FT_TJ), which the host checks.
template <bool F32>
__global__ void __launch_bounds__(FT_W / 2 * FT_NTY) k_mg_coarse_solve(LevelDev L, const double *__restrict__ f, double *__restrict__ x, int nsweeps) {
    IMPEC_DYN_SMEM(double, X);
    fused_body<FUSED_PLAIN, true, F32, FT_H>(L, f, (const double *)nullptr, x, 0, L.g, (const double *)nullptr, (double *)nullptr, (double *)nullptr, X, nsweeps, 0, 0, 1);
}

// ---------------------------------------------------------------- fused PCG search-direction update + mat-vec
//   d <- z + beta d   (FIRST: d <- z),   s <- S d,   part <- <s,d>       beta = scal[0]/scal[1] from device memory
// One pass over z, d and the links instead of k_update_d followed by k_matvec_dot (56 instead of 64 B/cell, one launch).
// Same tile scheme as k_mg_fused: the new d is staged in shared memory with a one-cell halo (recomputed redundantly).
template <bool FIRST, bool EDGE>
IMPEC_D void pcg_matvec_body_rw(const LevelDev &L, const double *__restrict__ scal, const double *__restrict__ z, const double *__restrict__ din,
                                double *__restrict__ d, double *__restrict__ s, double *__restrict__ part, double *__restrict__ X) {
    constexpr int HI = 2, TI = FT_H - 2 * HI;
    const Grid g = L.g;
    const int ld = g.ld;
    const int tx = threadIdx.x, ty = threadIdx.y & (FT_NTY - 1);
    const int jl = 2 * tx;
    const int j0 = (int)blockIdx.x * FT_TJ - FT_HJ, i0 = (int)blockIdx.y * TI - HI;
    const int jg = j0 + jl;
    const long long kbase = (long long)(i0 + ty + IMPEC_GH) * ld + (jg + IMPEC_PJ);
    const bool col_alloc = !EDGE || ((jg + IMPEC_PJ >= 0) && (jg + 1 + IMPEC_PJ < ld));
    const double beta = FIRST ? 0. : scal[0] / scal[1];
    const bool col_int = jl >= FT_HJ && jl < FT_W - FT_HJ && (!EDGE || jg < g.N2);
    // new search direction on the whole tile; the interior is written back
#pragma unroll
    for (int n = 0; n < FT_ROWS; n++) {
        const int r = ty + n * FT_NTY, ig = i0 + r;
        double a = 0., b = 0.;
        const bool row_alloc = !EDGE || ((ig >= -IMPEC_GH) && (ig < g.rows - IMPEC_GH));
        if (row_alloc && col_alloc) {
            const size_t k = (size_t)(kbase + (long long)n * FT_NTY * ld);
            const d2 zv = ld2(z + k);
            if (FIRST) { a = zv.x; b = zv.y; }
            else { const d2 dv = ld2(din + k); a = zv.x + beta * dv.x; b = zv.y + beta * dv.y; }
        }
        X[r * FT_W + tx] = a; X[r * FT_W + 64 + tx] = b;
    }
    IMPEC_SYNCTHREADS();
    double dotv = 0.;
#pragma unroll
    for (int n = 0; n < FT_ROWS; n++) {
        const int r = ty + n * FT_NTY, ig = i0 + r;
        if (r < HI || r >= FT_H - HI || !col_int || (EDGE && ig >= g.N1)) continue;
        const size_t k = (size_t)(kbase + (long long)n * FT_NTY * ld);
        const double da = X[r * FT_W + tx], db = X[r * FT_W + 64 + tx];
        const bool two = !EDGE || jg + 1 < g.N2;
        const Coef<false> A(L);
        double sa = is_pin(L, (long long)k) ? 0. : link_diag(L, k) * da - tile_offsum(A, ld, X, r, tx, 0, k);
        double sb = 0.;
        if (two) sb = is_pin(L, (long long)(k + 1)) ? 0. : link_diag(L, k + 1) * db - tile_offsum(A, ld, X, r, tx, 1, k + 1);
        if (two) { st2(d + k, da, db); st2(s + k, sa, sb); } else { d[k] = da; s[k] = sa; }
        dotv = fma(sa, da, dotv);
        if (two) dotv = fma(sb, db, dotv);
    }
    dotv = block_reduce<RED_SUM>(dotv);
    if (tx == 0 && ty == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = dotv;
}

// NOTE: d is read (with halo) and written (interior) by the same launch, so the read and write buffers must differ:
// the host ping-pongs between two search-direction buffers (din -> dout).
template <bool FIRST>
__global__ void __launch_bounds__(FT_W / 2 * FT_NTY, 2) k_pcg_matvec(LevelDev L, const double *__restrict__ scal, const double *__restrict__ z, const double *__restrict__ din,
                                                                 double *__restrict__ dout, double *__restrict__ s, double *__restrict__ part) {
    IMPEC_DYN_SMEM(double, X);
    constexpr int HI = 2, TI = FT_H - 2 * HI;
    const int j0 = (int)blockIdx.x * FT_TJ - FT_HJ, i0 = (int)blockIdx.y * TI - HI;
    bool edge = i0 < L.g.rlo || j0 < 0 || i0 + FT_H >= L.g.rhi || j0 + FT_W >= L.g.N2;
    if (L.pin >= 0) {
        const int pi = (int)(L.pin / L.g.ld) - IMPEC_GH, pj = (int)(L.pin % L.g.ld) - IMPEC_PJ, pje = pj + (int)(L.pin_end - L.pin);
        edge = edge || (pi >= i0 && pi < i0 + FT_H && pje > j0 && pj < j0 + FT_W);
    }
    if (edge) pcg_matvec_body_rw<FIRST, true>(L, scal, z, din, dout, s, part, X);
    else pcg_matvec_body_rw<FIRST, false>(L, scal, z, din, dout, s, part, X);
}

// ---------------------------------------------------------------- marching variant of the fused direction update + mat-vec
// Same operation as k_pcg_matvec, organised as a 2.5-D march instead of shared-memory tiles (the tile kernel spends most
// of its time with every warp waiting on the tile load: ncu long_scoreboard 25 warps per issue, 22 % issue activity, 57 %
// of DRAM peak).  A WARP owns 64 columns (a 128-bit column pair per lane) and marches down `nrows` rows:
//   * the new direction d = z + beta d of rows i-1, i, i+1 lives in registers; the two neighbour columns of a pair come
//     from the adjacent lanes by shuffle, and the two columns just outside the warp are recomputed by lanes 0 and 31
//     (two extra scalar loads per row and warp) -- no shared memory, no block barrier, warps never wait for each other;
//   * every link value is loaded ONCE per thread (rows i and i+1 of AW / AV are carried over to the next step), the tile
//     kernel fetched each twice through L1;
//   * the loads of the next MV_PF rows are in flight while a row is processed (~100 B per row and thread), so the
//     memory-level parallelism no longer depends on occupancy.
#define MV_BLOCK 128
#define MV_PF 2
struct MvRow { d2 z, d, aw, av, as; double ez, ed, av2, as2; };

template <bool FIRST>
IMPEC_D MvRow mv_load_row(const LevelDev &L, const double *__restrict__ z, const double *__restrict__ din, size_t k, bool ok, int eoff) {
    MvRow R;
    R.z.x = R.z.y = R.d.x = R.d.y = R.aw.x = R.aw.y = R.av.x = R.av.y = R.as.x = R.as.y = 0.;
    R.ez = R.ed = R.av2 = R.as2 = 0.;
    if (ok) {
        R.z = ld2(z + k);
        if (!FIRST) R.d = ld2(din + k);
        R.aw = ld2(L.AW + k); R.av = ld2(L.AV + k); R.as = ld2(L.AS + k);
        R.av2 = L.AV[k + 2]; R.as2 = L.AS[k + 2];
        if (eoff) { R.ez = z[(long long)k + eoff]; if (!FIRST) R.ed = din[(long long)k + eoff]; }
    }
    return R;
}

template <bool FIRST>
__global__ void __launch_bounds__(MV_BLOCK, 3) k_pcg_matvec_march(LevelDev L, const double *__restrict__ scal, const double *__restrict__ z, const double *__restrict__ din,
                                                                  double *__restrict__ dout, double *__restrict__ s, double *__restrict__ part, int nrows) {
    const Grid g = L.g;
    const size_t ld = (size_t)g.ld;
    const int lane = threadIdx.x & 31;
    const int jA = 2 * (int)(blockIdx.x * MV_BLOCK + threadIdx.x);       // columns jA, jA+1
    const int ia = chunk_of(L) * nrows;
    const int ib = (ia + nrows < g.N1) ? ia + nrows : g.N1;               // rows [ia, ib)
    const bool ok = jA < g.N2, two = jA + 1 < g.N2;
    // lane 0 also evaluates column jA-1, lane 31 column jA+2 (allocated: pad columns on both sides of a row)
    const int eoff = !ok ? 0 : (lane == 0 ? -1 : ((lane == 31 && jA + 2 <= g.N2) ? 2 : 0));
    const double beta = FIRST ? 0. : scal[0] / scal[1];
    const size_t kb = gidx(g, ia - 1, ok ? jA : 0);
    // state: direction on rows r-2, r-1, r (columns jA-1 | jA, jA+1 | jA+2) and the links of row r-1
    double l0 = 0., a0 = 0., b0 = 0., r0 = 0., l1 = 0., a1 = 0., b1 = 0., r1 = 0.;
    d2 awc, avc, asc; double avc2 = 0., asc2 = 0.;
    awc.x = awc.y = avc.x = avc.y = asc.x = asc.y = 0.;
    double dotv = 0.;
    MvRow q[MV_PF];
#pragma unroll
    for (int u = 0; u < MV_PF; u++) { const int row = (ia - 1 + u < ib) ? ia - 1 + u : ib; q[u] = mv_load_row<FIRST>(L, z, din, kb + (size_t)(row - ia + 1) * ld, ok, eoff); }
    for (int base = ia - 1; base <= ib; base += MV_PF) {
#pragma unroll
        for (int u = 0; u < MV_PF; u++) {
            const int row = base + u;
            if (row > ib) continue;
            const MvRow R = q[u];
            { const int nx = (row + MV_PF < ib) ? row + MV_PF : ib; q[u] = mv_load_row<FIRST>(L, z, din, kb + (size_t)(nx - ia + 1) * ld, ok, eoff); }
            if (L.pfd > 0 && row + L.pfd <= ib) {                        // L2 software prefetch of the row `pfd` steps ahead (see LevelDev::pfd)
                const size_t kp = kb + (size_t)(row + L.pfd - ia + 1) * ld;
                pf_l2(z + kp); if (!FIRST) pf_l2(din + kp); pf_l2(L.AW + kp); pf_l2(L.AV + kp); pf_l2(L.AS + kp);
            }
            // new direction on this row
            const double a2 = FIRST ? R.z.x : R.z.x + beta * R.d.x;
            const double b2 = FIRST ? R.z.y : R.z.y + beta * R.d.y;
            const double e2 = FIRST ? R.ez : R.ez + beta * R.ed;
            double l2 = IMPEC_SHFL_UP(b2, 1), r2 = IMPEC_SHFL_DOWN(a2, 1);
            if (lane == 0) l2 = e2;
            if (lane == 31) r2 = e2;
            if (row > ia) {                                              // s on row row-1 from rows row-2, row-1, row
                const size_t k = kb + (size_t)(row - ia) * ld;           // cell (row-1, jA)
                double dA = L.mass + (((awc.x + R.aw.x) + (asc.x + asc.y)) + ((avc.x + R.av.x) + (avc.y + R.av.y)));
                double dB = L.mass + (((awc.y + R.aw.y) + (asc.y + asc2)) + ((avc.y + R.av.y) + (avc2 + R.av2)));
                if (has_sig(L, (long long)k)) dA += L.sig[(long long)k - L.sig_idx];
                if (has_sig(L, (long long)(k + 1))) dB += L.sig[(long long)(k + 1) - L.sig_idx];
                double oA = awc.x * a0;
                oA = fma(R.aw.x, a2, oA); oA = fma(asc.x, l1, oA); oA = fma(asc.y, b1, oA);
                oA = fma(avc.x, l0, oA); oA = fma(R.av.x, l2, oA); oA = fma(avc.y, b0, oA); oA = fma(R.av.y, b2, oA);
                double oB = awc.y * b0;
                oB = fma(R.aw.y, b2, oB); oB = fma(asc.y, a1, oB); oB = fma(asc2, r1, oB);
                oB = fma(avc.y, a0, oB); oB = fma(R.av.y, a2, oB); oB = fma(avc2, r0, oB); oB = fma(R.av2, r2, oB);
                const double sA = is_pin(L, (long long)k) ? 0. : dA * a1 - oA;
                const double sB = is_pin(L, (long long)(k + 1)) ? 0. : dB * b1 - oB;
                if (two) { st2(dout + k, a1, b1); st2(s + k, sA, sB); dotv = fma(sA, a1, dotv); dotv = fma(sB, b1, dotv); }
                else if (ok) { dout[k] = a1; s[k] = sA; dotv = fma(sA, a1, dotv); }
            }
            l0 = l1; a0 = a1; b0 = b1; r0 = r1;
            l1 = l2; a1 = a2; b1 = b2; r1 = r2;
            awc = R.aw; avc = R.av; avc2 = R.av2; asc = R.as; asc2 = R.as2;
        }
    }
    dotv = block_reduce<RED_SUM>(dotv);
    if (threadIdx.x == 0) part[chunk_of(L) * gridDim.x + blockIdx.x] = dotv;
}
