// kernels_transport.inl -- ghost fills, source step, split advection sweeps, CFL reduction.
// Replaces the loop nests of the reference cited at each kernel (paths under /root/reference/).
// All arithmetic is FP64 and written in the reference's operation order; the library is compiled with
// -fmad=false so this explicit path is bit-identical to the CPU oracle.

struct DevParams {
    double dx, dy, x_lb, y_lb;
    double mu_res, mq;             // mq = (mu_res/mu_inj)^(1/4)
    double nu, porosity, inj_fluid_conc;
    double alpha_l, alpha_t;
    // rotated-frame constants (src/EllipticSolver.cpp:72-93)
    double dxi, deta, secamt2, cosa2, sina2, costsina, cosasint, cost2, sint2, cosa, sint, cost, sina;
    int avgproc, is_test, testno;
};

// ---------------------------------------------------------------- ghost fills (a1, part of a12)
IMPEC_HD int reflect_cell(int i, int N) { return i < 0 ? -i - 1 : (i >= N ? 2 * N - 1 - i : i); }
IMPEC_HD int wrap_cell(int i, int N) { return i < 0 ? i + N : (i >= N ? i - N : i); }

// Thread t enumerates the ghost cells only: first the 2*ng ghost rows (coalesced), then the 2*ng ghost
// columns of the interior rows.  Neumann: src/Util.cpp:283-324; periodic: src/Util.cpp:215-256.
__global__ void k_fill_scalar(Grid g, double *__restrict__ a, int periodic, int ng) {
    const long long W = g.N2 + 2 * ng;
    const long long nrowpart = 2LL * ng * W;
    const long long total = nrowpart + 2LL * ng * g.N1;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int i, j;
    if (t < nrowpart) {
        const int r = (int)(t / W);
        j = (int)(t % W) - ng;
        i = r < ng ? r - ng : g.N1 + (r - ng);
    } else {
        t -= nrowpart;
        i = (int)(t / (2 * ng));
        const int cidx = (int)(t % (2 * ng));
        j = cidx < ng ? cidx - ng : g.N2 + (cidx - ng);
    }
    if ((i < 0 && !g.wall_lo) || (i >= g.N1 && !g.wall_hi)) return;      // halo rows of a neighbouring slab: filled by the exchange
    const int si = periodic ? wrap_cell(i, g.N1) : reflect_cell(i, g.N1);
    const int sj = periodic ? wrap_cell(j, g.N2) : reflect_cell(j, g.N2);
    a[gidx(g, i, j)] = a[gidx(g, si, sj)];
}

// Closed form of the (order dependent) statement sequence in src/Util.cpp:331-438:
//   u(i,j) = s_i * u(face-mirror(i), cell-mirror(j)),  v(i,j) = s_j * v(cell-mirror(i), face-mirror(j)),
// with s = -1 when the normal index was mirrored.  u is filled on i in [-ng,N1+ng], j in [-ng,N2+ng);
// v on i in [-ng,N1+ng), j in [-ng,N2+ng].
__global__ void k_fill_velocity(Grid g, double *__restrict__ u, double *__restrict__ v, int ng) {
    const int W = g.N2 + 1 + 2 * ng, H = g.N1 + 1 + 2 * ng;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)W * H) return;
    const int i = (int)(t / W) - ng, j = (int)(t % W) - ng;
    if ((i < 0 && !g.wall_lo) || (i > g.N1 && !g.wall_hi)) return;       // halo rows of a neighbouring slab: filled by the exchange
    // u component
    if (j < g.N2 + ng && !(i >= 0 && i <= g.N1 && j >= 0 && j < g.N2)) {
        const int si = i < 0 ? -i : (i > g.N1 ? 2 * g.N1 - i : i);
        const int sj = reflect_cell(j, g.N2);
        const double val = u[gidx(g, si, sj)];
        u[gidx(g, i, j)] = (si != i) ? -val : val;
    }
    // v component
    if (i < g.N1 + ng && (i < g.N1 || g.wall_hi) && !(i >= 0 && i < g.N1 && j >= 0 && j <= g.N2)) {
        const int si = reflect_cell(i, g.N1);
        const int sj = j < 0 ? -j : (j > g.N2 ? 2 * g.N2 - j : j);
        const double val = v[gidx(g, si, sj)];
        v[gidx(g, i, j)] = (sj != j) ? -val : val;
    }
}

// ---------------------------------------------------------------- source step (a15)
// src/TRIntegrator.cpp:49-71.  q is non-zero only in the wells, so the step runs over a compact list
// of (cell, q) pairs built once at upload time instead of over the whole field.
__global__ void k_source_tr(int n, const int *__restrict__ cell_i, const int *__restrict__ cell_j,
                            const double *__restrict__ qv, Grid g, double *__restrict__ c,
                            double dt, double c_inj, double *__restrict__ massdiff) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double qij = qv[t];
    const size_t k = gidx(g, cell_i[t], cell_j[t]);
    double md = 0.0;
    if (qij < 0) {
        const double old = c[k];
        const double nw = old * (1 + 0.5 * dt * qij) / (1 - 0.5 * dt * qij);
        c[k] = nw;
        md = nw - old;
    } else if (qij > 0) {
        c[k] += dt * c_inj * qij;
    }
    massdiff[t] = md;
}

// ---------------------------------------------------------------- MUSCL-Hancock / upwind fluxes (a17, a18)
// src/HyperbolicSolver.cpp:208-210
template <bool PHI1> IMPEC_D double flux_fn(double vel, double conc, double phi) {
    return PHI1 ? vel * conc : vel * conc / phi;
}
// src/UpwindSplitSolver.cpp:63-65
template <bool PHI1> IMPEC_D double upwind_flux(double speed, double UL, double UR, double phi) {
    return (speed > 0) ? flux_fn<PHI1>(speed, UL, phi) : flux_fn<PHI1>(speed, UR, phi);
}
// src/MUSCLSplitSolver.cpp:61-85 (limiter argument ignored upstream: always the w = 0 van-Leer form)
IMPEC_D double muscl_slope(double du_m, double du_p) {
    double lim = 0.;
    if (du_m * du_p > 0) {
        double r = du_m / du_p;
        if (r > DBL_MAX) r = DBL_MAX;
        const double a = 2. * r / (1. + r);
        const double b = 2. / (1. + r);
        lim = (b < a) ? b : a;
    }
    return lim * 0.5 * (du_m + du_p);
}
// src/MUSCLSplitSolver.cpp:87-93 and :122-127: boundary-extrapolated values evolved by half a step
template <bool PHI1>
IMPEC_D void muscl_evolve(double vL, double vR, double Um, double U0, double Up, double h, double phi, double &ULbar, double &URbar) {
    const double d = muscl_slope(U0 - Um, Up - U0);
    const double UR = U0 + 0.5 * d, UL = U0 - 0.5 * d;
    const double df = h * (flux_fn<PHI1>(vL, UL, phi) - flux_fn<PHI1>(vR, UR, phi));
    URbar = UR + df;
    ULbar = UL + df;
}

// One directional sweep c_out = c_in + dt/h (F_lo - F_hi) with the fluxes evaluated on the fly (no F/G or
// U_Lin temporaries).  DIR 0: x sweep, stencil across rows (src/HyperbolicSplitSolver.cpp:113-133,
// src/MUSCLSplitSolver.cpp:102-141); DIR 1: y sweep, stencil inside a row (:139-159, :152-189).
// METHOD 0 MUSCL, 1 first-order upwind (src/UpwindSplitSolver.cpp:43-61).
// AVG: also write 0.5*(other + result) -- the symmetric-parallel average of
// src/HyperbolicSplitSolver.cpp:87-103 fused into the last sweep.
// The cell range [i0,i1) x [j0,j1) is chosen by the host (ghost strips are swept only where a later
// sweep reads them).
template <int DIR, int METHOD, bool PHI1, bool AVG>
__global__ void __launch_bounds__(256) k_sweep(Grid g, const double *__restrict__ cin, const double *__restrict__ vel,
                                               double *__restrict__ cout, const double *__restrict__ other,
                                               double dt, double hstep, double phi, int i0, int i1, int j0, int j1) {
    const int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= i1 || j >= j1) return;
    const int s = DIR == 0 ? g.ld : 1;                     // stride along the sweep direction
    const size_t k = gidx(g, i, j);
    const double c0 = cin[k];
    double Flo, Fhi;
    const double v0 = vel[k], v1 = vel[k + s];            // normal velocity on the low / high face of the cell
    if (METHOD == 1) {
        Flo = upwind_flux<PHI1>(v0, cin[k - s], c0, phi);
        Fhi = upwind_flux<PHI1>(v1, c0, cin[k + s], phi);
    } else {
        const double cm2 = cin[k - 2 * s], cm1 = cin[k - s], cp1 = cin[k + s], cp2 = cin[k + 2 * s];
        const double vm1 = vel[k - s], v2 = vel[k + 2 * s];
        const double h = 0.5 * dt / hstep;
        double Lm, Rm, L0, R0, Lp, Rp;
        muscl_evolve<PHI1>(vm1, v0, cm2, cm1, c0, h, phi, Lm, Rm);     // cell - 1
        muscl_evolve<PHI1>(v0, v1, cm1, c0, cp1, h, phi, L0, R0);      // this cell
        muscl_evolve<PHI1>(v1, v2, c0, cp1, cp2, h, phi, Lp, Rp);      // cell + 1
        Flo = upwind_flux<PHI1>(v0, Rm, L0, phi);
        Fhi = upwind_flux<PHI1>(v1, R0, Lp, phi);
    }
    double r = c0 + dt / hstep * (Flo - Fhi);
    if (AVG) r = 0.5 * (other[k] + r);
    cout[k] = r;
}

// Intercell fluxes as a field (CalculateIntercellFlux_F / _G of the reference's split solvers, src/MUSCLSplitSolver.cpp:102-189,
// src/UpwindSplitSolver.cpp:43-61).  The sweeps above never materialise F or G; this kernel exists for callers of the reference's public
// flux interface.  One thread per face (i, j) = the low face of cell (i, j) along DIR; same device functions as the sweeps.
template <int DIR, int METHOD, bool PHI1>
__global__ void __launch_bounds__(256) k_intercell_flux(Grid g, const double *__restrict__ cin, const double *__restrict__ vel, double *__restrict__ F,
                                                        double dt, double hstep, double phi, int i0, int i1, int j0, int j1) {
    const int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= i1 || j >= j1) return;
    const int s = DIR == 0 ? g.ld : 1;
    const size_t k = gidx(g, i, j);
    const double v0 = vel[k];
    if (METHOD == 1) { F[k] = upwind_flux<PHI1>(v0, cin[k - s], cin[k], phi); return; }
    const double h = 0.5 * dt / hstep;
    double Llo, Rlo, Lhi, Rhi;
    muscl_evolve<PHI1>(vel[k - s], v0, cin[k - 2 * s], cin[k - s], cin[k], h, phi, Llo, Rlo);       // cell on the low side of the face
    muscl_evolve<PHI1>(v0, vel[k + s], cin[k - s], cin[k], cin[k + s], h, phi, Lhi, Rhi);           // cell on the high side
    F[k] = upwind_flux<PHI1>(v0, Rlo, Lhi, phi);
}

// ---------------------------------------------------------------- MUSCL sweeps, one reconstruction per cell
// k_sweep evaluates the reconstruction of cells i-1, i, i+1 in every thread: 9 IEEE divisions per cell, which made the
// sweep FP64-pipe bound (ncu: 61 % fp64 pipe, 22 % of the HBM roofline).  The two kernels below evaluate every cell's
// reconstruction ONCE with the very same device functions (so results stay bit-identical to the oracle) and share it:
//
//  k_muscl_x: the stencil runs across rows.  A thread owns one column j and MARCHES down SX_ROWS rows keeping
//             c(i-1), c(i), c(i+1), the previous cell's evolved right state and the previous face flux in registers;
//             loads of the next four rows are issued before the arithmetic of the current four (coalesced along j).
//  k_muscl_y: the stencil runs inside a row.  A CTA owns a 256 x SY_ROWS tile, every thread reconstructs its own cells
//             and publishes the evolved states in shared memory; the 2*SY_ROWS halo cells of the tile are done by
//             the first lanes of warp 0.
#define SX_ROWS 64
#define SX_BLOCK 128
template <bool PHI1, bool AVG>
__global__ void __launch_bounds__(SX_BLOCK) k_muscl_x(Grid g, const double *__restrict__ cin, const double *__restrict__ vel,
                                                      double *__restrict__ cout, const double *__restrict__ other,
                                                      double dt, double hstep, double phi, int i0, int i1, int j0, int j1, int jstart, int pfd) {
    const int j = jstart + (int)(blockIdx.x * SX_BLOCK + threadIdx.x);
    if (j < j0 || j >= j1) return;
    const int ia = i0 + (int)blockIdx.y * SX_ROWS;
    const int ib = (ia + SX_ROWS < i1) ? ia + SX_ROWS : i1;          // cells [ia, ib) are written by this thread
    const size_t ld = (size_t)g.ld;
    const double h = 0.5 * dt / hstep;
    const size_t kb = gidx(g, ia - 1, j);                            // cell ia-1: reconstructed only
    double cm = cin[kb - ld], c0 = cin[kb], cp = cin[kb + ld];
    double v0 = vel[kb], v1 = vel[kb + ld];
    double L, Rprev, Fprev = 0.;
    muscl_evolve<PHI1>(v0, v1, cm, c0, cp, h, phi, L, Rprev);
    for (int ib0 = ia; ib0 <= ib; ib0 += 4) {
        double cn[4], vn[4], on[4];
        if (pfd > 0 && ib0 + pfd + 4 <= ib) {                        // the four rows `pfd` ahead into L2
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const size_t kp = kb + (size_t)(ib0 + pfd + u - ia + 2) * ld;
                pf_l2(cin + kp); pf_l2(vel + kp); if (AVG) pf_l2(other + kp - 2 * ld);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {                                // rows past ib are clamped: loaded, never used
            const int i = (ib0 + u < ib) ? ib0 + u : ib;
            const size_t k = kb + (size_t)(i - ia + 1) * ld;         // cell i
            cn[u] = cin[k + ld]; vn[u] = vel[k + ld];
            on[u] = AVG ? other[k - ld] : 0.;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = ib0 + u;
            if (i > ib) break;
            cm = c0; c0 = cp; cp = cn[u]; v0 = v1; v1 = vn[u];       // now centred on cell i
            double R;
            muscl_evolve<PHI1>(v0, v1, cm, c0, cp, h, phi, L, R);
            const double F = upwind_flux<PHI1>(v0, Rprev, L, phi);   // face i-1/2
            if (i > ia) {                                            // cell i-1 has both its fluxes
                double r = cm + dt / hstep * (Fprev - F);
                if (AVG) r = 0.5 * (on[u] + r);
                cout[kb + (size_t)(i - ia) * ld] = r;
            }
            Fprev = F; Rprev = R;
        }
    }
}

#define SY_ROWS 4
#define SY_BLOCK 256
template <bool PHI1, bool AVG>
__global__ void __launch_bounds__(SY_BLOCK) k_muscl_y(Grid g, const double *__restrict__ cin, const double *__restrict__ vel,
                                                      double *__restrict__ cout, const double *__restrict__ other,
                                                      double dt, double hstep, double phi, int i0, int i1, int j0, int j1) {
    __shared__ double Ls[SY_ROWS][SY_BLOCK + 2], Rs[SY_ROWS][SY_BLOCK + 2];
    const int tx = threadIdx.x;
    const int jb = j0 + (int)blockIdx.x * SY_BLOCK;                  // first cell of the tile
    const int jn = (j1 - jb < SY_BLOCK) ? j1 - jb : SY_BLOCK;        // cells per row in this tile
    const int ib = i0 + (int)blockIdx.y * SY_ROWS;
    const int rn = (i1 - ib < SY_ROWS) ? i1 - ib : SY_ROWS;
    const double h = 0.5 * dt / hstep;
    double cc[SY_ROWS], va[SY_ROWS], vb[SY_ROWS];
    if (tx < jn) {
#pragma unroll
        for (int n = 0; n < SY_ROWS; n++) {
            if (n >= rn) break;
            const size_t k = gidx(g, ib + n, jb + tx);
            const double cm = cin[k - 1], c0 = cin[k], cp = cin[k + 1];
            const double v0 = vel[k], v1 = vel[k + 1];
            double L, R;
            muscl_evolve<PHI1>(v0, v1, cm, c0, cp, h, phi, L, R);
            Ls[n][tx + 1] = L; Rs[n][tx + 1] = R;
            cc[n] = c0; va[n] = v0; vb[n] = v1;
        }
    }
    if (tx < 2 * SY_ROWS) {                                          // halo cells jb-1 and jb+jn of every row
        const int n = tx >> 1, side = tx & 1;
        if (n < rn) {
            const size_t k = gidx(g, ib + n, side ? jb + jn : jb - 1);
            double L, R;
            muscl_evolve<PHI1>(vel[k], vel[k + 1], cin[k - 1], cin[k], cin[k + 1], h, phi, L, R);
            Ls[n][side ? jn + 1 : 0] = L; Rs[n][side ? jn + 1 : 0] = R;
        }
    }
    IMPEC_SYNCTHREADS();
    if (tx < jn) {
#pragma unroll
        for (int n = 0; n < SY_ROWS; n++) {
            if (n >= rn) break;
            const size_t k = gidx(g, ib + n, jb + tx);
            const double Flo = upwind_flux<PHI1>(va[n], Rs[n][tx], Ls[n][tx + 1], phi);
            const double Fhi = upwind_flux<PHI1>(vb[n], Rs[n][tx + 1], Ls[n][tx + 2], phi);
            double r = cc[n] + dt / hstep * (Flo - Fhi);
            if (AVG) r = 0.5 * (other[k] + r);
            cout[k] = r;
        }
    }
}

// k_muscl_y2: the y sweep without shared memory or block barriers (the 256 x 4 tile kernel above sits at 39 % of the HBM roofline:
// 262144 tiny CTAs at 16384^2, a barrier and a serial halo pass each).  A warp owns 64 columns of a row -- one 128-bit column pair
// (cells A = jA, B = jA + 1) per lane --, every lane evolves its two cells once, the neighbours' values (c(jA-1), c(jA+2), the
// face velocity v(jA+2), the evolved right state of cell jA-1 and the evolved left state of cell jA+2) come from the adjacent lanes by
// shuffle.  The two outermost lanes lack a neighbour, so a warp stores 60 of its 64 columns (lanes 1..30).  Rows are independent in
// a y sweep: a thread walks down SY2_ROWS rows two at a time (two independent dependency chains), with the rows `pfd` ahead
// prefetched into L2.  Same device functions and operation order per cell as k_sweep / k_muscl_y: bit-identical.
#define SY2_BLOCK 128
#define SY2_OWN 60
#define SY2_ROWS 32
template <bool PHI1, bool AVG>
IMPEC_D void muscl_y2_row(const Grid &g, const double *__restrict__ cin, const double *__restrict__ vel, double *__restrict__ cout,
                          const double *__restrict__ other, size_t k, double dt, double hstep, double h, double phi, bool st2ok, bool st1ok) {
    const d2 c = ld2(cin + k), v = ld2(vel + k);
    const double cl = IMPEC_SHFL_UP(c.y, 1), cr = IMPEC_SHFL_DOWN(c.x, 1), vr = IMPEC_SHFL_DOWN(v.x, 1);
    double LA, RA, LB, RB;
    muscl_evolve<PHI1>(v.x, v.y, cl, c.x, c.y, h, phi, LA, RA);          // cell A
    muscl_evolve<PHI1>(v.y, vr, c.x, c.y, cr, h, phi, LB, RB);           // cell B
    const double Rl = IMPEC_SHFL_UP(RB, 1), Lr = IMPEC_SHFL_DOWN(LA, 1);  // right state of cell jA-1, left state of cell jA+2
    const double F0 = upwind_flux<PHI1>(v.x, Rl, LA, phi);
    const double F1 = upwind_flux<PHI1>(v.y, RA, LB, phi);
    const double F2 = upwind_flux<PHI1>(vr, RB, Lr, phi);
    double rA = c.x + dt / hstep * (F0 - F1);
    double rB = c.y + dt / hstep * (F1 - F2);
    if (st2ok) {
        if (AVG) { const d2 o = ld2(other + k); rA = 0.5 * (o.x + rA); rB = 0.5 * (o.y + rB); }
        st2(cout + k, rA, rB);
    } else if (st1ok) {
        if (AVG) rA = 0.5 * (other[k] + rA);
        cout[k] = rA;
    }
}
template <bool PHI1, bool AVG>
__global__ void __launch_bounds__(SY2_BLOCK) k_muscl_y2(Grid g, const double *__restrict__ cin, const double *__restrict__ vel,
                                                        double *__restrict__ cout, const double *__restrict__ other,
                                                        double dt, double hstep, double phi, int i0, int i1, int j0, int j1, int pfd) {
    const int lane = threadIdx.x & 31;
    const int gw = (int)blockIdx.x * (SY2_BLOCK / 32) + (int)(threadIdx.x >> 5);
    const int jA = j0 - 2 + SY2_OWN * gw + 2 * lane;                 // j0 is even (host): the pair is 16-byte aligned
    // columns jA, jA+1 are read when they lie in [j0 - 2, j1 + 2) (two ghost columns on either side are allocated and filled)
    const bool colok = jA + 1 < j1 + 2;
    const bool own = colok && lane >= 1 && lane < 31 && jA >= j0;
    const bool st2ok = own && jA + 1 < j1, st1ok = own && jA < j1;
    const int ia = i0 + (int)blockIdx.y * SY2_ROWS;
    const int ib = (ia + SY2_ROWS < i1) ? ia + SY2_ROWS : i1;
    const double h = 0.5 * dt / hstep;
    const size_t ld = (size_t)g.ld;
    const size_t kb = gidx(g, ia, colok ? jA : j0);
    for (int i = ia; i < ib; i += 2) {
        const size_t k = kb + (size_t)(i - ia) * ld;
        if (pfd > 0 && i + pfd + 1 < ib) {
            const size_t kp = k + (size_t)pfd * ld;
            pf_l2(cin + kp); pf_l2(vel + kp); pf_l2(cin + kp + ld); pf_l2(vel + kp + ld);
            if (AVG) { pf_l2(other + kp); pf_l2(other + kp + ld); }
        }
        muscl_y2_row<PHI1, AVG>(g, cin, vel, cout, other, k, dt, hstep, h, phi, st2ok, st1ok);
        // (an odd row count: the last iteration repeats row ib - 1 -- identical values stored twice -- so that all lanes of the warp stay
        //  on one path through the shuffles)
        muscl_y2_row<PHI1, AVG>(g, cin, vel, cout, other, (i + 1 < ib) ? k + ld : k, dt, hstep, h, phi, st2ok, st1ok);
    }
}

// ---------------------------------------------------------------- time-step reductions (a13)
// MODE 0: max over cells of max(|u_c|,|v_c|)            (src/HyperbolicSolver.cpp:123-156)
// MODE 1: min over cells of min(dx/|u_c|, dy/|v_c|)     (test mode, :100-121)
// MODE 2: min over cells of dx^2 / (2 (d11+d22))        (diffusive limit, :157-191)
template <int MODE>
__global__ void __launch_bounds__(256) k_dt_reduce(Grid g, const double *__restrict__ u, const double *__restrict__ v,
                                                   DevParams P, double *__restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double val = (MODE == 0) ? red_identity<RED_MAX>() : red_identity<RED_MIN>();
    if (i < g.N1 && j < g.N2) {
        const size_t k = gidx(g, i, j);
        if (MODE == 2) {
            const double vij = 0.5 * (v[k] + v[k + 1]);
            const double uij = 0.5 * (u[k] + u[k + g.ld]);
            const double v2 = vij * vij, u2 = uij * uij, absu = sqrt(v2 + u2);
            const double d11 = P.dx / absu * (P.alpha_l * u2 + P.alpha_t * v2);
            const double d22 = P.dx / absu * (P.alpha_l * v2 + P.alpha_t * u2);
            val = 1. / 2 * P.dx * P.dx / (d11 + d22);
        } else {
            double uc = fabs(0.5 * (u[k] + u[k + g.ld]));
            double vc = fabs(0.5 * (v[k] + v[k + 1]));
            if (uc < DBL_MIN) uc = DBL_MIN;
            if (vc < DBL_MIN) vc = DBL_MIN;
            if (MODE == 0) val = uc > vc ? uc : vc;
            else { const double a = P.dx / uc, b = P.dy / vc; val = b < a ? b : a; }
        }
    }
    val = (MODE == 0) ? block_reduce<RED_MAX>(val) : block_reduce<RED_MIN>(val);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = val;
}

// Stage 2 of every reduction: one block combines n partials in a fixed pairwise tree (deterministic).
// Up to 3 independent reductions are finished by one launch (blockIdx.x selects which).
template <int OP>
__global__ void __launch_bounds__(1024) k_reduce_final(const double *__restrict__ part, long long n, long long stride, double *__restrict__ out) {
    const double *p = part + (long long)blockIdx.x * stride;
    double acc = red_identity<OP>();
    // fixed assignment: thread t owns elements t, t+1024, ... combined in index order
    for (long long k = threadIdx.x; k < n; k += blockDim.x) acc = red_op<OP>(acc, p[k]);
    acc = block_reduce<OP>(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// ---------------------------------------------------------------- parallel -> diagonal projection
// ParallelProjector::FillDiagonal / GetFromParallel (src/ParallelProjector.cpp:91-199): one thread per cell of the diagonal
// grid; rotate its centre by -pi/4 into the parallel frame, mirror y, interpolate the ghosted parallel field bilinearly
// between the cells floor(x/L N) and +1 exactly as upstream (same operation order: bit-identical to the oracle).
// Slabs: `a` holds rows [i_off, i_off + g.N1) of a field of N1g rows (+ halo / ghost rows); a diagonal cell is evaluated by the rank that
// owns the lower row x00 of its stencil (the upper row is an own row or the first halo row), the other ranks write 0 and the chunk
// [row0, row0 + nrows) of the diagonal grid is summed over the ranks afterwards: one rank: own = everything, i_off = 0, N1g = g.N1.
struct ProjectArgs { double x_lb, y_lb, x_ub, y_ub, ca, sa, p_dx, p_dy, d_dx, d_dy; int dN1, dN2, ng; int N1g, i_off, own_lo, own_hi, row0, nrows; };
__global__ void __launch_bounds__(256) k_project_parallel(Grid g, const double *__restrict__ a, ProjectArgs A, double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int il = blockIdx.y * blockDim.y + threadIdx.y;
    const int i = A.row0 + il;
    if (il >= A.nrows || j >= A.dN2) return;
    const double xd = A.x_lb + (i + 0.5) * A.d_dx, yd = A.y_lb + (j + 0.5) * A.d_dy;
    const double x = A.ca * xd - A.sa * yd;
    const double y = fabs(A.sa * xd + A.ca * yd);
    int x00 = (int)floor(x / (A.x_ub - A.x_lb) * A.N1g), y00 = (int)floor(y / (A.y_ub - A.y_lb) * g.N2);
    // (upstream only warns about points outside the parallel domain; keep the reads inside the ghosted array)
    x00 = x00 < -A.ng ? -A.ng : (x00 > A.N1g + A.ng - 2 ? A.N1g + A.ng - 2 : x00);
    y00 = y00 < -A.ng ? -A.ng : (y00 > g.N2 + A.ng - 2 ? g.N2 + A.ng - 2 : y00);
    if (x00 < A.own_lo || x00 >= A.own_hi) { out[(size_t)il * A.dN2 + j] = 0.; return; }
    const double x_00 = A.x_lb + (x00 + 0.5) * A.p_dx, x_11 = A.x_lb + (x00 + 1 + 0.5) * A.p_dx;
    const double y_00 = A.y_lb + (y00 + 0.5) * A.p_dy, y_11 = A.y_lb + (y00 + 1 + 0.5) * A.p_dy;
    const size_t k = gidx(g, x00 - A.i_off, y00);
    out[(size_t)il * A.dN2 + j] =
        1. / ((x_11 - x_00) * (y_11 - y_00)) * (x_11 - x) * (y_11 - y) * a[k] +
        1. / ((x_11 - x_00) * (y_11 - y_00)) * (x - x_00) * (y_11 - y) * a[k + g.ld] +
        1. / ((x_11 - x_00) * (y_11 - y_00)) * (x_11 - x) * (y - y_00) * a[k + 1] +
        1. / ((x_11 - x_00) * (y_11 - y_00)) * (x - x_00) * (y - y_00) * a[k + g.ld + 1];
}
