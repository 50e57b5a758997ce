// kernels_march.inl -- the fused multigrid kernels as register MARCHES (big levels).
//
// Same mathematics, cell for cell and bit for bit, as k_mg_fused<PRE/POST/PLAIN> in kernels_fused.inl, but organised
// like k_pcg_matvec_march: a WARP owns 64 columns (one 128-bit column pair per lane) and marches down a chunk of rows.
// The iterate of the last few rows lives in registers, the neighbour columns of a pair come from the adjacent lanes by
// shuffle, nothing goes through shared memory and no warp ever waits at a block barrier.  The colour stages, the
// residual and the restriction are SKEWED along the march: when row t arrives,
//     PRE  : first colour on row t, second colour on row t-1, residual on row t-2, coarse row (t-4)/2;
//     POST : prolongation on row t, first colour on row t-1, second colour on row t-2 (stored, dotted with f).
// Halo: values that depend on a column outside the warp are wrong in the outermost lanes; the wrong region grows by one
// column per dependent stage, so PRE discards two lanes per side (56 of 64 columns are stored, the same 4-column halo as
// the tiles) and POST/PLAIN one lane per side (60 of 64).  Rows: a chunk of `nrows` rows starts 3 (PRE) / 2 (POST) rows
// early.  Each link value and each vector element is loaded once per warp that needs it, the loads of the next 2-4
// rows are in flight while a row is processed, and the row parity (which cell of the pair has which colour) is a
// template parameter of the step, so a step has no colour branches.
//
// Why: the tile kernels sit at 52-61 % of the HBM roofline with every warp stalled on the tile load (ncu long_scoreboard
// 8-13 warps per issue at 50 % occupancy) and 147 instructions per cell; the same reorganisation took the PCG mat-vec
// from 57 % to 95 % of the measured copy bandwidth.
#define MG_BLOCK 128
#define MG_PF_PRE 4              // PRE: rows in flight per thread (even: the row parity of a step is static); POST / PLAIN: 2
#define MGP_OWN 56                // columns a warp stores in PRE (lanes 2..29)
#define MGS_OWN 60                // columns a warp stores in POST / PLAIN (lanes 1..30)

template <bool F32> struct CoefType;
template <> struct CoefType<true> { typedef float T; };
template <> struct CoefType<false> { typedef double T; };
template <class T> struct cpair { T x, y; };
IMPEC_D cpair<float> ldc2(const float *p) {
#ifndef IMPEC_EMU
    const float2 v = *reinterpret_cast<const float2 *>(p);
    cpair<float> r; r.x = v.x; r.y = v.y; return r;
#else
    cpair<float> r; r.x = p[0]; r.y = p[1]; return r;
#endif
}
IMPEC_D cpair<double> ldc2(const double *p) { const d2 v = ld2(p); cpair<double> r; r.x = v.x; r.y = v.y; return r; }

// links and inverse diagonal of one row of a column pair: cell A uses .x (and .y for its right / upper-right link), cell B
// uses .y and the third value
template <bool F32> struct CoefRow {
    typedef typename CoefType<F32>::T T;
    cpair<T> dinv, aw, as, av; T as2, av2;
};
template <bool F32> IMPEC_D CoefRow<F32> coef_zero() {
    CoefRow<F32> c; c.dinv.x = c.dinv.y = c.aw.x = c.aw.y = c.as.x = c.as.y = c.av.x = c.av.y = 0; c.as2 = c.av2 = 0; return c;
}
template <bool F32> IMPEC_D CoefRow<F32> coef_load(const Coef<F32> &A, size_t k) {
    CoefRow<F32> c;
    c.dinv = ldc2(A.DINV + k); c.aw = ldc2(A.AW + k); c.as = ldc2(A.AS + k); c.av = ldc2(A.AV + k);
    c.as2 = A.AS[k + 2]; c.av2 = A.AV[k + 2];
    return c;
}
template <bool F32> IMPEC_D void coef_prefetch(const Coef<F32> &A, size_t k) { pf_l2(A.DINV + k); pf_l2(A.AW + k); pf_l2(A.AS + k); pf_l2(A.AV + k); }
// sum over the 8 neighbours in the order of tile_offsum: up, down, left, right, up-left, down-left, up-right, down-right
IMPEC_D double sum8(double w_up, double w_dn, double w_l, double w_r, double w_ul, double w_dl, double w_ur, double w_dr,
                    double up, double dn, double l, double r, double ul, double dl, double ur, double dr) {
    double s = w_up * up;
    s = fma(w_dn, dn, s); s = fma(w_l, l, s); s = fma(w_r, r, s);
    s = fma(w_ul, ul, s); s = fma(w_dl, dl, s); s = fma(w_ur, ur, s); s = fma(w_dr, dr, s);
    return s;
}
IMPEC_D double sum4(double w0, double w1, double w2, double w3, double v0, double v1, double v2, double v3) {
    double s = w0 * v0;
    s = fma(w1, v1, s); s = fma(w2, v2, s); s = fma(w3, v3, s);
    return s;
}

// ================================================================ PRE: sweep from zero + residual + restriction
template <bool F32> struct PreRow { d2 f; CoefRow<F32> c; };
// (k is a valid address for every lane: columns outside the level are clamped to column 0 by the caller and every use of the loaded
// values is masked by the act* / own flags, so the loads need no guard -- and no zero-fill of 12 registers per row)
template <bool F32> IMPEC_D PreRow<F32> pre_load(const Coef<F32> &A, const double *__restrict__ f, long long k, bool) {
    PreRow<F32> R;
    R.f = ld2(f + k); R.c = coef_load<F32>(A, (size_t)k);
    return R;
}
// Pipeline state in FOUR ROTATING SLOTS: the data of row r lives in slot r & 3 for as long as it is needed (links and right-hand side
// for 3 steps, final iterate values for 3 steps, the j-restricted residual for 4).  The march is unrolled four rows deep and the
// slot of every access is a compile-time constant of the unrolled step (template parameter PH = row & 3), so the state never moves:
// the earlier formulation shifted ~45 registers per row with IMAD.MOV (22 % of the kernel's instructions, ncu: issue-bound at 50 %).
template <bool F32> struct PreState {
    double xF[4], e0[4];                   // first-colour value of the row and the first-colour value next to its second-colour cell
    d2 f[4]; CoefRow<F32> c[4];
    double xl[4], xA[4], xB[4], xr[4];     // final values of the row: columns jA-1 | jA, jA+1 | jA+2
    double h[4];                           // j-restricted residual of the row
};
struct PreGeom {
    int t_rlo, t_rhi;                      // active rows
    int wlo, whi;                          // rows of x this chunk stores
    int Ilo, Ihi;                          // coarse rows this chunk stores
    int jA, J; bool colok, own, own_c, clamp_l, clamp_r, wall_lo, wall_hi;
    int N1;
    long long kb, ld, pin, pin_end;                 // index of (row 0, jA), pitch, pinned cell
    long long kcb, ldc;                    // coarse: index of (row 0, J), pitch
};

// AF: on row t the even-column cell A carries the first colour.  PH = t & 3 (slot of row t).
template <bool F32, bool AF, int PH>
IMPEC_D void pre_step(const int t, const PreRow<F32> &R, PreState<F32> &S, const PreGeom &G, double *__restrict__ x, double *__restrict__ fc) {
    constexpr int s0 = PH, s1 = (PH + 3) & 3, s2 = (PH + 2) & 3, s3 = (PH + 1) & 3;      // slots of rows t, t-1, t-2, t-3
    // 1. first colour on row t (zero guess: no neighbours)
    const bool act0 = G.colok && t >= G.t_rlo && t < G.t_rhi;
    const double xF = act0 ? (AF ? R.f.x * (double)R.c.dinv.x : R.f.y * (double)R.c.dinv.y) : 0.;
    const double e0 = AF ? IMPEC_SHFL_DOWN(xF, 1) : IMPEC_SHFL_UP(xF, 1);
    // 2. second colour on row t-1: its face neighbours carry final first-colour values, its diagonal neighbours are still zero
    const bool act1 = G.colok && t - 1 >= G.t_rlo && t - 1 < G.t_rhi;
    double xl1, xA1, xB1, xr1;
    if (AF) {   // row t-1: B first, A second
        double off = (double)S.c[s1].aw.x * S.xA[s2];
        off = fma((double)R.c.aw.x, xF, off); off = fma((double)S.c[s1].as.x, S.e0[s1], off); off = fma((double)S.c[s1].as.y, S.xF[s1], off);
        const double xS = act1 ? (S.f[s1].x + off) * (double)S.c[s1].dinv.x : 0.;
        xA1 = xS; xB1 = S.xF[s1]; xl1 = S.e0[s1]; xr1 = IMPEC_SHFL_DOWN(xS, 1);
    } else {    // row t-1: A first, B second
        double off = (double)S.c[s1].aw.y * S.xB[s2];
        off = fma((double)R.c.aw.y, xF, off); off = fma((double)S.c[s1].as.y, S.xF[s1], off); off = fma((double)S.c[s1].as2, S.e0[s1], off);
        const double xS = act1 ? (S.f[s1].y + off) * (double)S.c[s1].dinv.y : 0.;
        xA1 = S.xF[s1]; xB1 = xS; xr1 = S.e0[s1]; xl1 = IMPEC_SHFL_UP(xS, 1);
    }
    if (G.own && t - 1 >= G.wlo && t - 1 < G.whi) st2(x + G.kb + (long long)(t - 1) * G.ld, xA1, xB1);
    // 3. residual on row t-2 (same parity as row t).  After a sweep from zero: first colour = sum over all 8 neighbours,
    //    second colour = sum over its 4 (same-colour) diagonal neighbours
    const bool act2 = G.colok && t - 2 >= G.t_rlo && t - 2 < G.t_rhi;
    const long long k2 = G.kb + (long long)(t - 2) * G.ld;
    double rA, rB;
    if (AF) {
        rA = sum8((double)S.c[s2].aw.x, (double)S.c[s1].aw.x, (double)S.c[s2].as.x, (double)S.c[s2].as.y, (double)S.c[s2].av.x, (double)S.c[s1].av.x, (double)S.c[s2].av.y, (double)S.c[s1].av.y,
                  S.xA[s3], xA1, S.xl[s2], S.xB[s2], S.xl[s3], xl1, S.xB[s3], xB1);
        rB = sum4((double)S.c[s2].av.y, (double)S.c[s1].av.y, (double)S.c[s2].av2, (double)S.c[s1].av2, S.xA[s3], xA1, S.xr[s3], xr1);
    } else {
        rB = sum8((double)S.c[s2].aw.y, (double)S.c[s1].aw.y, (double)S.c[s2].as.y, (double)S.c[s2].as2, (double)S.c[s2].av.y, (double)S.c[s1].av.y, (double)S.c[s2].av2, (double)S.c[s1].av2,
                  S.xB[s3], xB1, S.xA[s2], S.xr[s2], S.xA[s3], xA1, S.xr[s3], xr1);
        rA = sum4((double)S.c[s2].av.x, (double)S.c[s1].av.x, (double)S.c[s2].av.y, (double)S.c[s1].av.y, S.xl[s3], xl1, S.xB[s3], xB1);
    }
    if (!act2 || (k2 >= G.pin && k2 < G.pin_end)) rA = 0.;
    if (!act2 || (k2 + 1 >= G.pin && k2 + 1 < G.pin_end)) rB = 0.;
    // 4. full weighting: along j now, along i when the fourth row of a coarse cell is complete
    double rl = IMPEC_SHFL_UP(rB, 1), rr = IMPEC_SHFL_DOWN(rA, 1);
    if (G.clamp_l) rl = rA;
    if (G.clamp_r) rr = rB;
    double h = 0.25 * rl;
    h += 0.75 * rA; h += 0.75 * rB; h += 0.25 * rr;
    S.h[s2] = h;                                   // rows t-5, t-4, t-3, t-2 of h sit in slots s1, s0, s3, s2
    if (AF) {                                      // t even: rows t-5 .. t-2 are rows 2I-1 .. 2I+2 of coarse row I
        const int I = (t - 4) >> 1;
        if (G.own_c && I >= G.Ilo && I < G.Ihi) {
            const double t0 = (I == 0 && G.wall_lo) ? S.h[s0] : S.h[s1];                  // row -1 mirrors onto row 0
            const double t3 = (2 * I + 2 > G.N1 - 1 && G.wall_hi) ? S.h[s3] : S.h[s2];   // row N1 mirrors onto row N1-1
            fc[G.kcb + (long long)I * G.ldc] = 0.25 * (0.25 * t0 + 0.75 * S.h[s0] + 0.75 * S.h[s3] + 0.25 * t3);
        }
    }
    // the row's data stay where they are: only the new values are written into the slots of rows t and t-1
    S.xl[s1] = xl1; S.xA[s1] = xA1; S.xB[s1] = xB1; S.xr[s1] = xr1;
    S.xF[s0] = xF; S.e0[s0] = e0; S.f[s0] = R.f; S.c[s0] = R.c;
}

template <bool F32, int PF, int MINB>
__global__ void __launch_bounds__(MG_BLOCK, MINB) k_mg_march_pre(LevelDev L, const double *__restrict__ f, double *__restrict__ x, Grid gc, double *__restrict__ fc, int nrows) {
    static_assert(PF == 2 || PF == 4, "rows in flight");
    const Grid g = L.g;
    const Coef<F32> A(L);
    const int lane = threadIdx.x & 31;
    const int gw = (int)blockIdx.x * (MG_BLOCK / 32) + (int)(threadIdx.x >> 5);
    const int ia = chunk_of(L) * nrows;        // nrows is a multiple of 4: the slot of a row is a constant of the unrolled step
    const int ib = (ia + nrows < g.N1) ? ia + nrows : g.N1;
    PreGeom G;
    G.jA = MGP_OWN * gw - 4 + 2 * lane; G.J = G.jA >> 1;
    G.colok = G.jA >= 0 && G.jA < g.N2;
    G.own = G.colok && lane >= 2 && lane < 30;
    G.own_c = G.own && G.J < gc.N2;
    G.clamp_l = G.jA == 0; G.clamp_r = G.jA + 2 > g.N2 - 1;
    G.wall_lo = g.wall_lo != 0; G.wall_hi = g.wall_hi != 0; G.N1 = g.N1;
    G.t_rlo = g.rlo; G.t_rhi = g.rhi;
    // a slab's first / last chunk also stores x on the 2 halo rows it has computed anyway (saves an exchange, see k_mg_fused)
    G.wlo = (ia == 0 && !g.wall_lo) ? -2 : ia;
    G.whi = (ib == g.N1 && !g.wall_hi) ? g.N1 + 2 : ib;
    G.Ilo = ia >> 1; G.Ihi = (ib >> 1 < gc.N1) ? ib >> 1 : gc.N1;
    G.ld = g.ld; G.kb = (long long)IMPEC_GH * g.ld + (G.colok ? G.jA : 0) + IMPEC_PJ; G.pin = L.pin; G.pin_end = L.pin_end;
    G.ldc = gc.ld; G.kcb = (long long)IMPEC_GH * gc.ld + (G.colok ? G.J : 0) + IMPEC_PJ;
    PreState<F32> S;
#pragma unroll
    for (int a = 0; a < 4; a++) {
        S.xF[a] = S.e0[a] = 0.; S.f[a].x = S.f[a].y = 0.; S.c[a] = coef_zero<F32>();
        S.xl[a] = S.xA[a] = S.xB[a] = S.xr[a] = 0.; S.h[a] = 0.;
    }
    const int t0 = ia - 3, t1 = ib + 2;        // rows marched: [t0, t1]; t0 = 1 (mod 4)
    PreRow<F32> q[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) q[u] = pre_load<F32>(A, f, G.kb + (long long)((t0 + u <= t1) ? t0 + u : t1) * G.ld, G.colok);
    for (int t = t0; t <= t1; t += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (t + u > t1) continue;
            const PreRow<F32> R = q[u % PF];
            q[u % PF] = pre_load<F32>(A, f, G.kb + (long long)((t + u + PF <= t1) ? t + u + PF : t1) * G.ld, G.colok);
            if (L.pfd > 0 && t + u + L.pfd <= t1) { const long long kp = G.kb + (long long)(t + u + L.pfd) * G.ld; pf_l2(f + kp); coef_prefetch<F32>(A, (size_t)kp); }
            // row t + u: slot (1 + u) & 3; A carries the first colour on even rows (u odd)
            if (u == 0) pre_step<F32, false, 1>(t, R, S, G, x, fc);
            else if (u == 1) pre_step<F32, true, 2>(t + 1, R, S, G, x, fc);
            else if (u == 2) pre_step<F32, false, 3>(t + 2, R, S, G, x, fc);
            else pre_step<F32, true, 0>(t + 3, R, S, G, x, fc);
        }
    }
}

// ================================================================ POST / PLAIN: (prolongation +) sweep from the current iterate
template <bool F32, bool PROLONG> struct PostRow { d2 xin, f; CoefRow<F32> c; double e[6]; };
template <bool F32> struct PostState {
    double il1, iA1, iB1, ir1, il2, iA2, iB2, ir2, il3, iA3, iB3, ir3;   // iterate before the sweep on rows t-1, t-2, t-3
    double yF2, yF3, e2;                   // new first-colour values of rows t-2, t-3; first-colour value next to row t-2's second-colour cell
    d2 f1, f2; CoefRow<F32> c1, c2;        // rows t-1, t-2
};
struct PostGeom {
    int t_rlo, t_rhi, ia, ib;
    int jA; bool colok, own;
    long long kb, ld, pin, pin_end;
};
template <bool F32, bool PROLONG>
IMPEC_D PostRow<F32, PROLONG> post_load(const Coef<F32> &A, const double *__restrict__ f, const double *__restrict__ xin, const Grid &g, const Grid &gc,
                                         const double *__restrict__ ec, const PostGeom &G, int t) {
    PostRow<F32, PROLONG> R;
    R.xin.x = R.xin.y = R.f.x = R.f.y = 0.; R.c = coef_zero<F32>();
#pragma unroll
    for (int a = 0; a < 6; a++) R.e[a] = 0.;
    if (G.colok) {
        const long long k = G.kb + (long long)t * G.ld;
        R.xin = ld2(xin + k); R.f = ld2(f + k); R.c = coef_load<F32>(A, (size_t)k);
        if (PROLONG && t >= G.t_rlo && t < G.t_rhi) {
            // cell-centred bilinear prolongation with edge replication (same clamps as k_mg_fused)
            const int I = t >> 1, J = G.jA >> 1;
            int In = (t & 1) ? I + 1 : I - 1, Jm = J - 1, Jp = J + 1;
            if (In < 0 && gc.wall_lo) In = 0;
            if (In > gc.N1 - 1 && gc.wall_hi) In = gc.N1 - 1;
            Jm = Jm < 0 ? 0 : Jm; Jp = Jp > gc.N2 - 1 ? gc.N2 - 1 : Jp;
            const double *e0 = ec + gidx(gc, I, 0), *e1 = ec + gidx(gc, In, 0);
            R.e[0] = e0[J]; R.e[1] = e1[J]; R.e[2] = e0[Jm]; R.e[3] = e1[Jm]; R.e[4] = e0[Jp]; R.e[5] = e1[Jp];
        }
    }
    return R;
}

// AF: on row t the even-column cell A carries the first colour (so do rows t-2; rows t-1, t-3 have B first)
template <bool F32, bool PROLONG, bool AF>
IMPEC_D void post_step(const int t, const PostRow<F32, PROLONG> &R, PostState<F32> &S, const PostGeom &G, double *__restrict__ x, double &dotv) {
    // 1. iterate on row t: x + P ec
    double a = R.xin.x, b = R.xin.y;
    if (PROLONG && G.colok && t >= G.t_rlo && t < G.t_rhi) {
        const double pc = 0.75 * R.e[0] + 0.25 * R.e[1];
        const double pm = 0.75 * R.e[2] + 0.25 * R.e[3];
        const double pp = 0.75 * R.e[4] + 0.25 * R.e[5];
        const long long k = G.kb + (long long)t * G.ld;
        if (!(k >= G.pin && k < G.pin_end)) a += 0.75 * pc + 0.25 * pm;
        if (!(k + 1 >= G.pin && k + 1 < G.pin_end)) b += 0.75 * pc + 0.25 * pp;
    }
    const double il0 = IMPEC_SHFL_UP(b, 1), ir0 = IMPEC_SHFL_DOWN(a, 1);
    // 2. first colour on row t-1 (rows t-2, t-1, t of the old iterate; same-colour diagonal neighbours lagged)
    const bool act1 = G.colok && t - 1 >= G.t_rlo && t - 1 < G.t_rhi;
    double y, e1;
    if (AF) {   // row t-1: B first
        const double off = sum8((double)S.c1.aw.y, (double)R.c.aw.y, (double)S.c1.as.y, (double)S.c1.as2, (double)S.c1.av.y, (double)R.c.av.y, (double)S.c1.av2, (double)R.c.av2,
                                S.iB2, b, S.iA1, S.ir1, S.iA2, a, S.ir2, ir0);
        y = act1 ? (S.f1.y + off) * (double)S.c1.dinv.y : S.iB1;
        e1 = IMPEC_SHFL_UP(y, 1);
    } else {    // row t-1: A first
        const double off = sum8((double)S.c1.aw.x, (double)R.c.aw.x, (double)S.c1.as.x, (double)S.c1.as.y, (double)S.c1.av.x, (double)R.c.av.x, (double)S.c1.av.y, (double)R.c.av.y,
                                S.iA2, a, S.il1, S.iB1, S.il2, il0, S.iB2, b);
        y = act1 ? (S.f1.x + off) * (double)S.c1.dinv.x : S.iA1;
        e1 = IMPEC_SHFL_DOWN(y, 1);
    }
    // 3. second colour on row t-2: face neighbours new, diagonal neighbours (its own colour) old
    const bool act2 = G.colok && t - 2 >= G.t_rlo && t - 2 < G.t_rhi;
    double xA, xB;
    if (AF) {   // row t-2: A first (new value yF2), B second
        const double off = sum8((double)S.c2.aw.y, (double)S.c1.aw.y, (double)S.c2.as.y, (double)S.c2.as2, (double)S.c2.av.y, (double)S.c1.av.y, (double)S.c2.av2, (double)S.c1.av2,
                                S.yF3, y, S.yF2, S.e2, S.iA3, S.iA1, S.ir3, S.ir1);
        xA = S.yF2; xB = act2 ? (S.f2.y + off) * (double)S.c2.dinv.y : S.iB2;
    } else {    // row t-2: B first, A second
        const double off = sum8((double)S.c2.aw.x, (double)S.c1.aw.x, (double)S.c2.as.x, (double)S.c2.as.y, (double)S.c2.av.x, (double)S.c1.av.x, (double)S.c2.av.y, (double)S.c1.av.y,
                                S.yF3, y, S.e2, S.yF2, S.il3, S.il1, S.iB3, S.iB1);
        xB = S.yF2; xA = act2 ? (S.f2.x + off) * (double)S.c2.dinv.x : S.iA2;
    }
    if (G.own && t - 2 >= G.ia && t - 2 < G.ib) {
        st2(x + G.kb + (long long)(t - 2) * G.ld, xA, xB);
        dotv = fma(S.f2.x, xA, dotv); dotv = fma(S.f2.y, xB, dotv);
    }
    // shift the pipeline
    S.il3 = S.il2; S.iA3 = S.iA2; S.iB3 = S.iB2; S.ir3 = S.ir2;
    S.il2 = S.il1; S.iA2 = S.iA1; S.iB2 = S.iB1; S.ir2 = S.ir1;
    S.il1 = il0; S.iA1 = a; S.iB1 = b; S.ir1 = ir0;
    S.yF3 = S.yF2; S.yF2 = y; S.e2 = e1;
    S.f2 = S.f1; S.f1 = R.f; S.c2 = S.c1; S.c1 = R.c;
}

// C0: colour of the first stage (0 forward, 1 backward); part: per-CTA partial sums of <f, x> (or nullptr)
template <bool F32, bool PROLONG, int C0, int MINB>
__global__ void __launch_bounds__(MG_BLOCK, MINB) k_mg_march_post(LevelDev L, const double *__restrict__ f, const double *__restrict__ xin, double *__restrict__ x,
                                                               Grid gc, const double *__restrict__ ec, double *__restrict__ part, int nrows) {
    const Grid g = L.g;
    const Coef<F32> A(L);
    const int lane = threadIdx.x & 31;
    const int gw = (int)blockIdx.x * (MG_BLOCK / 32) + (int)(threadIdx.x >> 5);
    PostGeom G;
    G.ia = chunk_of(L) * nrows;
    G.ib = (G.ia + nrows < g.N1) ? G.ia + nrows : g.N1;
    G.jA = MGS_OWN * gw - 2 + 2 * lane;
    G.colok = G.jA >= 0 && G.jA < g.N2;
    G.own = G.colok && lane >= 1 && lane < 31;
    G.t_rlo = g.rlo; G.t_rhi = g.rhi;
    G.ld = g.ld; G.kb = (long long)IMPEC_GH * g.ld + (G.colok ? G.jA : 0) + IMPEC_PJ; G.pin = L.pin; G.pin_end = L.pin_end;
    PostState<F32> S;
    S.il1 = S.iA1 = S.iB1 = S.ir1 = S.il2 = S.iA2 = S.iB2 = S.ir2 = S.il3 = S.iA3 = S.iB3 = S.ir3 = 0.;
    S.yF2 = S.yF3 = S.e2 = 0.; S.f1.x = S.f1.y = S.f2.x = S.f2.y = 0.; S.c1 = coef_zero<F32>(); S.c2 = coef_zero<F32>();
    double dotv = 0.;
    const int t0 = G.ia - 2, t1 = G.ib + 1;    // rows marched: [t0, t1], an even count; t0 is even
    PostRow<F32, PROLONG> q0 = post_load<F32, PROLONG>(A, f, xin, g, gc, ec, G, t0), q1 = post_load<F32, PROLONG>(A, f, xin, g, gc, ec, G, t0 + 1);
    for (int t = t0; t <= t1; t += 2) {
        if (L.pfd > 0 && t + 1 + L.pfd <= t1) {          // rows t + pfd and t + pfd + 1 (and the coarse row under them) into L2
            const long long kp = G.kb + (long long)(t + L.pfd) * G.ld;
            pf_l2(xin + kp); pf_l2(f + kp); coef_prefetch<F32>(A, (size_t)kp);
            pf_l2(xin + kp + G.ld); pf_l2(f + kp + G.ld); coef_prefetch<F32>(A, (size_t)(kp + G.ld));
            if (PROLONG) { const int I = (t + L.pfd) >> 1; if (I >= 0 && I < gc.N1 + 1) pf_l2(ec + (long long)(IMPEC_GH + I) * gc.ld + (G.colok ? (G.jA >> 1) : 0) + IMPEC_PJ); }
        }
        { const PostRow<F32, PROLONG> R = q0; q0 = post_load<F32, PROLONG>(A, f, xin, g, gc, ec, G, (t + 2 <= t1) ? t + 2 : t1);
          post_step<F32, PROLONG, C0 == 0>(t, R, S, G, x, dotv); }
        { const PostRow<F32, PROLONG> R = q1; q1 = post_load<F32, PROLONG>(A, f, xin, g, gc, ec, G, (t + 3 <= t1) ? t + 3 : t1);
          post_step<F32, PROLONG, C0 != 0>(t + 1, R, S, G, x, dotv); }
    }
    if (part) {
        dotv = block_reduce<RED_SUM>(dotv);
        if (threadIdx.x == 0) part[chunk_of(L) * gridDim.x + blockIdx.x] = dotv;
    }
}

// ================================================================ symmetric Gauss-Seidel from zero in ONE march
// The Crank-Nicolson dispersion solve (a19) preconditions PCG with one symmetric red-black sweep from a zero guess.  As launches
// that is a memset of x, a forward sweep (k_mg_march_post<PLAIN, 0>) and a backward sweep (<PLAIN, 1>): 8 + 40 + 40 B/cell, and a
// separate <r,z> pass.  Here the four colour stages run skewed along one march -- when row t arrives: colour 0 on row t (zero
// guess: f * dinv), colour 1 on row t-1 (forward sweep done on that row), colour 1 again on row t-2 (first stage of the backward
// sweep, all eight neighbours from the forward result) and colour 0 on row t-3 (stored, dotted with f): 34 B/cell in one launch.
// The arithmetic per cell is that of the three launches (the forward half is pre_step's, the backward half post_step's), so the
// result is bit-identical; two halo lanes per side as in PRE.
template <bool F32> struct SgsFwd { double xF1, e01, xA2, xB2; d2 f1; CoefRow<F32> c1; };
template <bool F32, bool AF>
IMPEC_D void sgs_step(const int t, const PreRow<F32> &R, SgsFwd<F32> &F, PostState<F32> &S, const PostGeom &G, double *__restrict__ x, double &dotv) {
    // ---- forward sweep from zero (pre_step, parts 1 and 2): colour 0 on row t, colour 1 on row t-1
    const bool act0 = G.colok && t >= G.t_rlo && t < G.t_rhi;
    const double xF = act0 ? (AF ? R.f.x * (double)R.c.dinv.x : R.f.y * (double)R.c.dinv.y) : 0.;
    const double e0 = AF ? IMPEC_SHFL_DOWN(xF, 1) : IMPEC_SHFL_UP(xF, 1);
    const bool actf = G.colok && t - 1 >= G.t_rlo && t - 1 < G.t_rhi;
    double il0, a, b, ir0;                                     // forward result on row t-1: columns jA-1 | jA, jA+1 | jA+2
    if (AF) {
        double off = (double)F.c1.aw.x * F.xA2;
        off = fma((double)R.c.aw.x, xF, off); off = fma((double)F.c1.as.x, F.e01, off); off = fma((double)F.c1.as.y, F.xF1, off);
        const double xS = actf ? (F.f1.x + off) * (double)F.c1.dinv.x : 0.;
        a = xS; b = F.xF1; il0 = F.e01; ir0 = IMPEC_SHFL_DOWN(xS, 1);
    } else {
        double off = (double)F.c1.aw.y * F.xB2;
        off = fma((double)R.c.aw.y, xF, off); off = fma((double)F.c1.as.y, F.xF1, off); off = fma((double)F.c1.as2, F.e01, off);
        const double xS = actf ? (F.f1.y + off) * (double)F.c1.dinv.y : 0.;
        a = F.xF1; b = xS; ir0 = F.e01; il0 = IMPEC_SHFL_UP(xS, 1);
    }
    // ---- backward sweep (post_step, parts 2 and 3, one row later): its "row t" is our row t-1, whose links are F.c1.
    //      First colour of the backward sweep = colour 1, carried by A on row t-1 exactly when A carries colour 0 on row t.
    const int tb = t - 1;
    const bool act1 = G.colok && tb - 1 >= G.t_rlo && tb - 1 < G.t_rhi;
    double y, e1;
    if (AF) {   // row tb-1: B first
        const double off = sum8((double)S.c1.aw.y, (double)F.c1.aw.y, (double)S.c1.as.y, (double)S.c1.as2, (double)S.c1.av.y, (double)F.c1.av.y, (double)S.c1.av2, (double)F.c1.av2,
                                S.iB2, b, S.iA1, S.ir1, S.iA2, a, S.ir2, ir0);
        y = act1 ? (S.f1.y + off) * (double)S.c1.dinv.y : S.iB1;
        e1 = IMPEC_SHFL_UP(y, 1);
    } else {    // row tb-1: A first
        const double off = sum8((double)S.c1.aw.x, (double)F.c1.aw.x, (double)S.c1.as.x, (double)S.c1.as.y, (double)S.c1.av.x, (double)F.c1.av.x, (double)S.c1.av.y, (double)F.c1.av.y,
                                S.iA2, a, S.il1, S.iB1, S.il2, il0, S.iB2, b);
        y = act1 ? (S.f1.x + off) * (double)S.c1.dinv.x : S.iA1;
        e1 = IMPEC_SHFL_DOWN(y, 1);
    }
    const bool act2 = G.colok && tb - 2 >= G.t_rlo && tb - 2 < G.t_rhi;
    double xA, xB;
    if (AF) {   // row tb-2: A first (new value yF2), B second
        const double off = sum8((double)S.c2.aw.y, (double)S.c1.aw.y, (double)S.c2.as.y, (double)S.c2.as2, (double)S.c2.av.y, (double)S.c1.av.y, (double)S.c2.av2, (double)S.c1.av2,
                                S.yF3, y, S.yF2, S.e2, S.iA3, S.iA1, S.ir3, S.ir1);
        xA = S.yF2; xB = act2 ? (S.f2.y + off) * (double)S.c2.dinv.y : S.iB2;
    } else {    // row tb-2: B first, A second
        const double off = sum8((double)S.c2.aw.x, (double)S.c1.aw.x, (double)S.c2.as.x, (double)S.c2.as.y, (double)S.c2.av.x, (double)S.c1.av.x, (double)S.c2.av.y, (double)S.c1.av.y,
                                S.yF3, y, S.e2, S.yF2, S.il3, S.il1, S.iB3, S.iB1);
        xB = S.yF2; xA = act2 ? (S.f2.x + off) * (double)S.c2.dinv.x : S.iA2;
    }
    if (G.own && tb - 2 >= G.ia && tb - 2 < G.ib) {
        st2(x + G.kb + (long long)(tb - 2) * G.ld, xA, xB);
        dotv = fma(S.f2.x, xA, dotv); dotv = fma(S.f2.y, xB, dotv);
    }
    // shift both pipelines
    S.il3 = S.il2; S.iA3 = S.iA2; S.iB3 = S.iB2; S.ir3 = S.ir2;
    S.il2 = S.il1; S.iA2 = S.iA1; S.iB2 = S.iB1; S.ir2 = S.ir1;
    S.il1 = il0; S.iA1 = a; S.iB1 = b; S.ir1 = ir0;
    S.yF3 = S.yF2; S.yF2 = y; S.e2 = e1;
    S.f2 = S.f1; S.f1 = F.f1; S.c2 = S.c1; S.c1 = F.c1;
    F.xA2 = a; F.xB2 = b; F.xF1 = xF; F.e01 = e0; F.f1 = R.f; F.c1 = R.c;
}

template <bool F32>
__global__ void __launch_bounds__(MG_BLOCK, 3) k_sgs_march(LevelDev L, const double *__restrict__ f, double *__restrict__ x, double *__restrict__ part, int nrows) {
    const Grid g = L.g;
    const Coef<F32> A(L);
    const int lane = threadIdx.x & 31;
    const int gw = (int)blockIdx.x * (MG_BLOCK / 32) + (int)(threadIdx.x >> 5);
    PostGeom G;
    G.ia = chunk_of(L) * nrows;
    G.ib = (G.ia + nrows < g.N1) ? G.ia + nrows : g.N1;
    G.jA = MGP_OWN * gw - 4 + 2 * lane;
    G.colok = G.jA >= 0 && G.jA < g.N2;
    G.own = G.colok && lane >= 2 && lane < 30;
    G.t_rlo = g.rlo; G.t_rhi = g.rhi;
    G.ld = g.ld; G.kb = (long long)IMPEC_GH * g.ld + (G.colok ? G.jA : 0) + IMPEC_PJ; G.pin = L.pin; G.pin_end = L.pin_end;
    SgsFwd<F32> F;
    F.xF1 = F.e01 = F.xA2 = F.xB2 = 0.; F.f1.x = F.f1.y = 0.; F.c1 = coef_zero<F32>();
    PostState<F32> S;
    S.il1 = S.iA1 = S.iB1 = S.ir1 = S.il2 = S.iA2 = S.iB2 = S.ir2 = S.il3 = S.iA3 = S.iB3 = S.ir3 = 0.;
    S.yF2 = S.yF3 = S.e2 = 0.; S.f1.x = S.f1.y = S.f2.x = S.f2.y = 0.; S.c1 = coef_zero<F32>(); S.c2 = coef_zero<F32>();
    double dotv = 0.;
    const int t0 = G.ia - 3, t1 = G.ib + 2;    // rows marched: [t0, t1], an even count; t0 is odd
    PreRow<F32> q0 = pre_load<F32>(A, f, G.kb + (long long)t0 * G.ld, G.colok), q1 = pre_load<F32>(A, f, G.kb + (long long)(t0 + 1) * G.ld, G.colok);
    for (int t = t0; t <= t1; t += 2) {
        if (L.pfd > 0 && t + 1 + L.pfd <= t1) {
            const long long kp = G.kb + (long long)(t + L.pfd) * G.ld;
            pf_l2(f + kp); coef_prefetch<F32>(A, (size_t)kp); pf_l2(f + kp + G.ld); coef_prefetch<F32>(A, (size_t)(kp + G.ld));
        }
        { const PreRow<F32> R = q0; q0 = pre_load<F32>(A, f, G.kb + (long long)((t + 2 <= t1) ? t + 2 : t1) * G.ld, G.colok);
          sgs_step<F32, false>(t, R, F, S, G, x, dotv); }
        { const PreRow<F32> R = q1; q1 = pre_load<F32>(A, f, G.kb + (long long)((t + 3 <= t1) ? t + 3 : t1) * G.ld, G.colok);
          sgs_step<F32, true>(t + 1, R, F, S, G, x, dotv); }
    }
    if (part) {
        dotv = block_reduce<RED_SUM>(dotv);
        if (threadIdx.x == 0) part[chunk_of(L) * gridDim.x + blockIdx.x] = dotv;
    }
}
