// kernels_build.inl -- the pressure operator of one step in ONE pass over the concentration (production runs: K = 1).
//
// First generation (kernels_pressure.inl, still used by the manufactured-solution test mode): k_mobility (c -> MU), k_vertex
// (MU -> KV), k_links (MU, KV -> AW, AS, AV), k_diag_inv (links -> DINV), k_links_to_f32 (4 arrays -> FP32 copies): five passes,
// 168 B per cell, 12.1 ms of every 16384^2 step (profiles/r1q_launches_summary.txt).  Here a WARP owns 64 columns (one 128-bit
// column pair per lane, 60 stored) and marches down a chunk of rows; the viscosities of the last two rows, the vertex mobilities,
// the rotated-frame cross terms and the links of the previous row live in registers, the neighbour columns come from the adjacent
// lanes by shuffle.  Row t's x-face and vertex links are complete when row t of c has arrived; the y-face links and the inverse
// diagonal of row t-1 need the vertex row t and are emitted one row later.  Traffic: R c (8) + W AW, AS, AV (24) + W the FP32 copies
// of AW, AS, AV, DINV (16) [+ W DINV in FP64 (8) where something still reads it] = 48 (56) B per cell.
//
// Every expression is the one of k_mobility / k_vertex / k_links / k_diag_inv, operand for operand (a2-a8: src/EllipticSolver.cpp:478-524,
// 692-754; src/MGSolver.cpp:710-841, folding :305-515), so the arrays are BIT-IDENTICAL to the five-pass build
// (tests: check_fused_build).
#define BL_BLOCK 128
#define BL_OWN 60                 // columns a warp stores (lanes 1..30)
#define BL_ROWS 64                // rows of a chunk

struct BuildArgs {
    Grid g;
    const double *c;
    double *AW, *AS, *AV, *DINV;
    float *AWf, *ASf, *AVf, *DINVf;
    long long pin, pin_end;
    DevParams P;
};

IMPEC_D void st2f(float *p, float a, float b) {
#ifndef IMPEC_EMU
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}

// vertex mobility from the four viscosities around vertex (i-1/2, j-1/2), in k_vertex's operand order
IMPEC_D double bl_vertex(int avgproc, double mu_ij, double mu_ip1j, double mu_ijp1, double mu_ip1jp1) {
    return (avgproc == 1) ? 4. / (mu_ip1j + mu_ij + mu_ijp1 + mu_ip1jp1) : 1. / 4 * (1. / mu_ip1j + 1. / mu_ij + 1. / mu_ijp1 + 1. / mu_ip1jp1);
}

template <bool W64>
__global__ void __launch_bounds__(BL_BLOCK) k_build_links_march(BuildArgs A, int nrows) {
    const Grid g = A.g; const DevParams &P = A.P;
    const int lane = threadIdx.x & 31;
    const int warp = (int)blockIdx.x * (BL_BLOCK / 32) + (threadIdx.x >> 5);
    const int ja = warp * BL_OWN - 2 + 2 * lane, jb = ja + 1;             // this lane's column pair
    const int i0 = (int)blockIdx.y * nrows;
    int i1 = i0 + nrows; if (i1 >= g.N1) i1 = g.N1 + 1;                   // the chunk that reaches N1 also owns the face/vertex row N1
    const bool own = lane >= 1 && lane <= 30 && ja < g.N2;
    const bool loadok = ja >= -2 && jb <= g.N2 + 1;                       // inside the ghosted row of c
    const long long ld = g.ld;
    const long long kb = (long long)gidx(g, 0, 0) + ja;                   // index of (row 0, ja)
    const double nu = P.nu, one_m_nu = 1 - nu;
    const double inv_xi2 = P.dxi * P.dxi, inv_xe = 1. / (P.deta * P.dxi);
    const double rdx2 = 1. / (P.dx * P.dx), rdy2 = 1. / (P.dy * P.dy);
    const bool rot = (nu != 1.);
    const int avg = P.avgproc;
    // state of row t-1
    double mu1A = 0., mu1B = 0., mu1L = 0.;
    double m121a = 0., m121b = 0., avr1a = 0., avr1b = 0.;               // M12 and raw vertex link of vertex row t-1
    double aw1a = 0., aw1b = 0., av1a = 0., av1b = 0., av1R = 0.;        // stored links of row t-1
    d2 cn = {0., 0.}, cn2 = {0., 0.};                                     // rows t+1 and t+2 of c (in flight)
    if (loadok) {
        cn = ld2(A.c + kb + (long long)(i0 - 1) * ld);
        if (i0 <= g.N1) cn2 = ld2(A.c + kb + (long long)i0 * ld);
    }
    for (int t = i0 - 1; t <= i1 && t <= g.N1; t++) {
        const d2 cc = cn; cn = cn2;
        if (loadok && t + 2 <= g.N1 && t + 2 <= i1) cn2 = ld2(A.c + kb + (long long)(t + 2) * ld);
        // viscosity of row t (k_mobility)
        const double mu0A = mu_of_c(cc.x, P.mq, P.mu_res), mu0B = mu_of_c(cc.y, P.mq, P.mu_res);
        const double mu0L = IMPEC_SHFL_UP(mu0B, 1);
        // vertex row t (k_vertex) and what k_links derives from it
        const double kva = bl_vertex(avg, mu1L, mu0L, mu1A, mu0A), kvb = bl_vertex(avg, mu1A, mu0A, mu1B, mu0B);
        const double m120a = vM12(P, kva, kva), m120b = vM12(P, kvb, kvb);
        const double avr0a = one_m_nu * (vM11(P, kva, kva) / inv_xi2), avr0b = one_m_nu * (vM11(P, kvb, kvb) / inv_xi2);
        const double m120R = IMPEC_SHFL_DOWN(m120a, 1), avr0R = IMPEC_SHFL_DOWN(avr0a, 1);
        const bool at_lo = (t == 0 && g.wall_lo), at_hi = (t == g.N1 && g.wall_hi);
        const bool xwall = at_lo || at_hi;
        const double av0a = (rot && !(xwall || ja == 0 || ja == g.N2)) ? avr0a : 0.;
        const double av0b = (rot && !(xwall || jb == 0 || jb == g.N2)) ? avr0b : 0.;
        const double av0R = IMPEC_SHFL_DOWN(av0a, 1);
        // x-face links of row t (between rows t-1 and t)
        double aw0a = 0., aw0b = 0.;
        if (!xwall) {
            if (ja < g.N2) {
                const double Wp = rdx2 * ((avg == 1) ? 2. / (mu1A + mu0A) : 1. / 2 * (1. / mu1A + 1. / mu0A));
                double Wd = 0.;
                if (rot) Wd = -inv_xe * (m120b + m120a);
                double w = nu * Wp + one_m_nu * Wd;
                if (rot && ja == 0) w = w + avr0a;
                if (rot && ja == g.N2 - 1) w = w + avr0b;
                aw0a = w;
            }
            if (jb < g.N2) {
                const double Wp = rdx2 * ((avg == 1) ? 2. / (mu1B + mu0B) : 1. / 2 * (1. / mu1B + 1. / mu0B));
                double Wd = 0.;
                if (rot) Wd = -inv_xe * (m120R + m120b);
                double w = nu * Wp + one_m_nu * Wd;
                if (rot && jb == g.N2 - 1) w = w + avr0R;
                aw0b = w;
            }
        }
        const bool st0 = own && t >= i0 && t < i1 && t >= 0;
        if (st0) {
            const long long k = kb + (long long)t * ld;
            st2(A.AW + k, aw0a, aw0b); st2(A.AV + k, av0a, av0b);
            st2f(A.AWf + k, (float)aw0a, (float)aw0b); st2f(A.AVf + k, (float)av0a, (float)av0b);
        }
        // y-face links and inverse diagonal of row t-1
        {
            const int i = t - 1;
            double asa = 0., asb = 0.;
            if (i >= 0 && i < g.N1) {
                if (ja > 0 && ja < g.N2) {
                    const double Sp = rdy2 * ((avg == 1) ? 2. / (mu1L + mu1A) : 1. / 2 * (1. / mu1L + 1. / mu1A));
                    double Sd = 0.;
                    if (rot) Sd = inv_xe * (m121a + m120a);
                    double s = nu * Sp + one_m_nu * Sd;
                    if (rot && i == 0 && g.wall_lo) s = s + avr1a;
                    if (rot && i == g.N1 - 1 && g.wall_hi) s = s + avr0a;
                    asa = s;
                }
                if (jb > 0 && jb < g.N2) {
                    const double Sp = rdy2 * ((avg == 1) ? 2. / (mu1A + mu1B) : 1. / 2 * (1. / mu1A + 1. / mu1B));
                    double Sd = 0.;
                    if (rot) Sd = inv_xe * (m121b + m120b);
                    double s = nu * Sp + one_m_nu * Sd;
                    if (rot && i == 0 && g.wall_lo) s = s + avr1b;
                    if (rot && i == g.N1 - 1 && g.wall_hi) s = s + avr0b;
                    asb = s;
                }
            }
            const double asR = IMPEC_SHFL_DOWN(asa, 1);
            if (own && i >= i0 && i < i1 && i >= 0 && i < g.N1) {
                const long long k = kb + (long long)i * ld;
                const double da = (((aw1a + aw0a) + (asa + asb)) + ((av1a + av0a) + (av1b + av0b)));
                const double db = (((aw1b + aw0b) + (asb + asR)) + ((av1b + av0b) + (av1R + av0R)));
                const double dia = (k >= A.pin && k < A.pin_end) ? 0. : 1. / da;
                const double dib = (jb < g.N2) ? ((k + 1 >= A.pin && k + 1 < A.pin_end) ? 0. : 1. / db) : 0.;
                st2(A.AS + k, asa, asb); st2f(A.ASf + k, (float)asa, (float)asb);
                st2f(A.DINVf + k, (float)dia, (float)dib);
                if (W64) st2(A.DINV + k, dia, dib);
            }
        }
        mu1A = mu0A; mu1B = mu0B; mu1L = mu0L;
        m121a = m120a; m121b = m120b; avr1a = avr0a; avr1b = avr0b;
        aw1a = aw0a; aw1b = aw0b; av1a = av0a; av1b = av0b; av1R = av0R;
    }
}

// DINV (interior) and the FP32 copies of all four coefficient arrays of one level in one pass (k_diag_inv + k_links_to_f32):
// coarse levels, the manufactured-solution test mode and the Crank-Nicolson operator.  The FP32 copies cover ALL allocated rows
// (halo rows included); outside the interior DINVf is the copy of what DINV holds (halo rows of a slab arrive by exchange).
template <bool W64>
__global__ void __launch_bounds__(256) k_level_finish(LevelDev L, const double *__restrict__ sigma, double *__restrict__ DINV,
                                                      float *__restrict__ AWf, float *__restrict__ ASf, float *__restrict__ AVf, float *__restrict__ DINVf) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y * blockDim.y + threadIdx.y;
    if (row >= L.g.rows || col >= L.g.ld) return;
    const size_t k = (size_t)row * L.g.ld + col;
    const int i = row - IMPEC_GH, j = col - IMPEC_PJ;
    const double aw = L.AW[k], as = L.AS[k], av = L.AV[k];
    AWf[k] = (float)aw; ASf[k] = (float)as; AVf[k] = (float)av;
    if (i >= 0 && i < L.g.N1 && j >= 0 && j < L.g.N2) {
        const int ld = L.g.ld;
        double d = L.mass + (((aw + L.AW[k + ld]) + (as + L.AS[k + 1])) + ((av + L.AV[k + ld]) + (L.AV[k + 1] + L.AV[k + ld + 1])));
        if (has_sig(L, (long long)k) && sigma) d += sigma[(long long)k - L.sig_idx];
        const double di = is_pin(L, (long long)k) ? 0. : 1. / d;
        if (W64) DINV[k] = di;
        DINVf[k] = (float)di;
    } else
        DINVf[k] = W64 ? (float)DINV[k] : 0.f;
}

// FP32 copies of rows [r0, r1) of the four arrays (the halo rows of a slab after their exchange)
__global__ void __launch_bounds__(256) k_f32_rows(Grid g, int r0, int r1, const double *__restrict__ AW, const double *__restrict__ AS, const double *__restrict__ AV,
                                                  const double *__restrict__ DINV, float *__restrict__ AWf, float *__restrict__ ASf, float *__restrict__ AVf, float *__restrict__ DINVf) {
    const long long n = (long long)(r1 - r0) * g.ld;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long k = (long long)r0 * g.ld + t;
    AWf[k] = (float)AW[k]; ASf[k] = (float)AS[k]; AVf[k] = (float)AV[k]; DINVf[k] = (float)DINV[k];
}

// ================================================================ a12 + a13: Darcy velocity and the CFL reductions in one march
// First generation: k_mobility + k_vertex (c -> MU, KV: 32 B/cell), k_velocity (R p, MU, KV, W u, v: 40 B/cell, every vertex velocity
// evaluated three times), k_dt_reduce (R u, v: 16 B/cell, twice with dispersion).  Here: the same column-pair march as the operator
// build, from c and p directly.  Row t's vertex velocities are evaluated ONCE and give u(t,.) at once and v(t-1,.) one row later;
// the cell-centred speeds of row t-1 (max |u|, |v| for the advective limit, the dispersion-tensor limit when DIFF) are reduced on the
// fly.  Traffic: R c, p (16) + W u, v (16) = 32 B/cell.  Expressions as in k_velocity / k_dt_reduce (src/EllipticSolver.cpp:541-683,
// src/HyperbolicSolver.cpp:123-196), operand for operand: u, v BIT-IDENTICAL, and max / min do not depend on the order.
struct VelArgs {
    Grid g;
    const double *c, *p;
    double *u, *v;
    double *part; long long stride;     // per-CTA partials: [0] max speed, [stride] min diffusive limit
    DevParams P;
};

IMPEC_D void bl_vertex_vel(const DevParams &P, double kv, double p_k, double p_kmldm1, double p_kmld, double p_km1, double &u1, double &u2) {
    const double d1dp = 1. / P.dxi * (p_k - p_kmldm1);
    const double d2dp = 1. / P.deta * (p_kmld - p_km1);
    const double M11 = vM11(P, kv, kv), M12 = vM12(P, kv, kv), M22 = vM22(P, kv, kv);
    const double w1 = -d1dp * M11 - d2dp * M12;
    const double w2 = -d1dp * M12 - d2dp * M22;
    u1 = P.cost * w1 - P.sina * w2;
    u2 = P.sint * w1 + P.cosa * w2;
}

template <bool DIFF>
__global__ void __launch_bounds__(BL_BLOCK) k_velocity_march(VelArgs A, int nrows) {
    const Grid g = A.g; const DevParams &P = A.P;
    const int lane = threadIdx.x & 31;
    const int warp = (int)blockIdx.x * (BL_BLOCK / 32) + (threadIdx.x >> 5);
    const int ja = warp * BL_OWN - 2 + 2 * lane, jb = ja + 1;
    const int i0 = (int)blockIdx.y * nrows;
    int i1 = i0 + nrows; if (i1 >= g.N1) i1 = g.N1 + 1;                   // the chunk that reaches N1 also owns the face row N1 of u
    const bool own = lane >= 1 && lane <= 30 && ja <= g.N2;
    const bool loadok = ja >= -2 && jb <= g.N2 + 1;
    const long long ld = g.ld;
    const long long kb = (long long)gidx(g, 0, 0) + ja;
    const double nu = P.nu, one_m_nu = 1. - nu;
    const bool rot = (nu != 1.);
    const int avg = P.avgproc;
    double mu1A = 0., mu1B = 0., mu1L = 0., p1A = 0., p1B = 0., p1L = 0.;
    double u21a = 0., u21b = 0.;                 // u2 of vertex row t-1
    double u1a_prev = 0., u1b_prev = 0.;         // u(t-1, .) (stored values) for the cell-centred speed
    double vmax = red_identity<RED_MAX>(), dmin = red_identity<RED_MIN>();
    d2 cn = {0., 0.}, pn = {0., 0.};
    if (loadok) { cn = ld2(A.c + kb + (long long)(i0 - 1) * ld); pn = ld2(A.p + kb + (long long)(i0 - 1) * ld); }
    for (int t = i0 - 1; t <= i1 && t <= g.N1; t++) {
        const d2 cc = cn, pp = pn;
        if (loadok && t + 1 <= g.N1 && t + 1 <= i1) { cn = ld2(A.c + kb + (long long)(t + 1) * ld); pn = ld2(A.p + kb + (long long)(t + 1) * ld); }
        const double mu0A = mu_of_c(cc.x, P.mq, P.mu_res), mu0B = mu_of_c(cc.y, P.mq, P.mu_res);
        const double mu0L = IMPEC_SHFL_UP(mu0B, 1);
        const double p0A = pp.x, p0B = pp.y, p0L = IMPEC_SHFL_UP(p0B, 1);
        // vertex velocities of vertex row t
        double u10a = 0., u20a = 0., u10b = 0., u20b = 0.;
        if (rot) {
            const double kva = bl_vertex(avg, mu1L, mu0L, mu1A, mu0A), kvb = bl_vertex(avg, mu1A, mu0A, mu1B, mu0B);
            bl_vertex_vel(P, kva, p0A, p1L, p1A, p0L, u10a, u20a);
            bl_vertex_vel(P, kvb, p0B, p1A, p1B, p0A, u10b, u20b);
        }
        const double u10R = IMPEC_SHFL_DOWN(u10a, 1);
        // u on the x-faces of row t
        const double fa = (avg == 1) ? 2. / (mu1A + mu0A) : 1. / 2 * (1. / mu1A + 1. / mu0A);
        const double fb = (avg == 1) ? 2. / (mu1B + mu0B) : 1. / 2 * (1. / mu1B + 1. / mu0B);
        const double ua = nu * (-fa * (p0A - p1A) / P.dx) + one_m_nu * (0.5 * (u10b + u10a));
        const double ub = nu * (-fb * (p0B - p1B) / P.dx) + one_m_nu * (0.5 * (u10R + u10b));
        if (own && t >= i0 && t < i1 && t >= 0) st2(A.u + kb + (long long)t * ld, ua, ub);
        // v on the y-faces of row t-1, cell-centred speeds of row t-1
        {
            const int i = t - 1;
            const double ga = (avg == 1) ? 2. / (mu1L + mu1A) : 1. / 2 * (1. / mu1L + 1. / mu1A);
            const double gb = (avg == 1) ? 2. / (mu1A + mu1B) : 1. / 2 * (1. / mu1A + 1. / mu1B);
            const double va = nu * (-ga * (p1A - p1L) / P.dy) + one_m_nu * (0.5 * (u21a + u20a));
            const double vb = nu * (-gb * (p1B - p1A) / P.dy) + one_m_nu * (0.5 * (u21b + u20b));
            const double vR = IMPEC_SHFL_DOWN(va, 1);
            if (own && i >= i0 && i < i1 && i >= 0 && i < g.N1) {
                st2(A.v + kb + (long long)i * ld, va, vb);
#define BL_SPEED(UPREV, UCUR, VL, VR, OK)                                                         \
                if (OK) {                                                                         \
                    if (DIFF) {                                                                   \
                        const double vij = 0.5 * ((VL) + (VR)), uij = 0.5 * ((UPREV) + (UCUR));   \
                        const double v2 = vij * vij, u2 = uij * uij, absu = sqrt(v2 + u2);        \
                        const double d11 = P.dx / absu * (P.alpha_l * u2 + P.alpha_t * v2);       \
                        const double d22 = P.dx / absu * (P.alpha_l * v2 + P.alpha_t * u2);       \
                        const double lim = 1. / 2 * P.dx * P.dx / (d11 + d22);                    \
                        dmin = lim < dmin ? lim : dmin;                                           \
                    }                                                                             \
                    double uc = fabs(0.5 * ((UPREV) + (UCUR))), vc = fabs(0.5 * ((VL) + (VR)));   \
                    if (uc < DBL_MIN) uc = DBL_MIN;                                               \
                    if (vc < DBL_MIN) vc = DBL_MIN;                                               \
                    const double m = uc > vc ? uc : vc;                                           \
                    vmax = m > vmax ? m : vmax;                                                   \
                }
                BL_SPEED(u1a_prev, ua, va, vb, ja < g.N2)
                BL_SPEED(u1b_prev, ub, vb, vR, jb < g.N2)
#undef BL_SPEED
            }
        }
        mu1A = mu0A; mu1B = mu0B; mu1L = mu0L; p1A = p0A; p1B = p0B; p1L = p0L;
        u21a = u20a; u21b = u20b; u1a_prev = ua; u1b_prev = ub;
    }
    vmax = block_reduce<RED_MAX>(vmax);
    if (DIFF) dmin = block_reduce<RED_MIN>(dmin);
    if (threadIdx.x == 0) {
        const long long b = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        A.part[b] = vmax;
        if (DIFF) A.part[A.stride + b] = dmin;
    }
}
