// kernels_pressure.inl -- mobility/link build, multigrid (smoother, residual, transfers), PCG vector
// kernels and Darcy velocity reconstruction.
//
// Operator representation ("links").  After the reference's Neumann folding (src/MGSolver.cpp:305-515)
// the 9-point row of cell (i,j) is   (A p)(i,j) = sum_nb a_nb (p_nb - p_ij)   with symmetric weights:
//   AW(i,j): (i-1,j)-(i,j)      AS(i,j): (i,j-1)-(i,j)
//   AV(i,j): lives on vertex (i-1/2,j-1/2) and couples (i-1,j-1)-(i,j) and (i-1,j)-(i,j-1)
// and every link that crosses the domain boundary is ZERO (its weight was folded onto the mirror link).
// Ghost values of all multigrid vectors are kept finite (zero) so that kernels need no edge branches.
// We solve the SPD form  S = -A :  (S x) = diag*x - sum a_nb x_nb,  diag = mass + sigma + sum a_nb.

// ---------------------------------------------------------------- a2-a4: viscosity and mobilities
// src/EllipticSolver.cpp:692-695
IMPEC_D double mu_of_c(double c, double mq, double mu_res) {
    const double x = (1. - c + mq * c);
    return 1. / (x * x * x * x) * mu_res;
}
// src/EllipticSolver.cpp:418-470 (manufactured-solution permeabilities; 1 in production runs)
IMPEC_D double K11_fn(const DevParams &P, double x, double y) {
    switch (P.testno) {
    case 2: return exp(-0.3 * x);
    case 3: return 2. + sin(2. * x) * sin(y);
    case 4: return 2. + sin(2. * x) * sin(y);
    default: return 1.;
    }
}
IMPEC_D double K22_fn(const DevParams &P, double x, double y) {
    switch (P.testno) {
    case 2: return exp(-0.3 * y);
    case 3: return 2. + sin(2. * x) * sin(y);
    case 4: return 2. + sin(x) * sin(2. * y);
    default: return 1.;
    }
}

// MU(i,j) = mu(c(i,j)) on the interior plus one ghost ring (c ghosts must be filled).
__global__ void __launch_bounds__(256) k_mobility(Grid g, const double *__restrict__ c, double *__restrict__ MU, double mq, double mu_res) {
    const int j = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 1;
    const int i = (int)(blockIdx.y * blockDim.y + threadIdx.y) - 1;
    if (i > g.N1 || j > g.N2) return;
    const size_t k = gidx(g, i, j);
    MU[k] = mu_of_c(c[k], mq, mu_res);
}

// face mobility between cell a (lower index) and cell b: src/EllipticSolver.cpp:478-524
template <bool TEST> IMPEC_D double face_mob(double mu_a, double mu_b, double Ka, double Kb, int avgproc) {
    if (avgproc == 1) return TEST ? 2. / (mu_a / Ka + mu_b / Kb) : 2. / (mu_a + mu_b);
    return TEST ? 1. / 2 * (Ka / mu_a + Kb / mu_b) : 1. / 2 * (1. / mu_a + 1. / mu_b);
}

// Vertex mobilities K11v,K22v at vertex (i-1/2,j-1/2), i in [0,N1], j in [0,N2]: src/EllipticSolver.cpp:697-754.
// In production K = 1 so K11v == K22v bit for bit and only KV11 is written.
template <bool TEST>
__global__ void __launch_bounds__(256) k_vertex(Grid g, const double *__restrict__ MU, double *__restrict__ KV11, double *__restrict__ KV22, DevParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > g.N1 || j > g.N2) return;
    const size_t k = gidx(g, i, j);
    // the reference's argument order: (i0,j0)=(i-1,j-1): c_ij, c_ip1j, c_ijp1, c_ip1jp1, summed as ip1j, ij, ijp1, ip1jp1
    const double mu_ij = MU[k - g.ld - 1], mu_ip1j = MU[k - 1], mu_ijp1 = MU[k - g.ld], mu_ip1jp1 = MU[k];
    if (!TEST) {
        KV11[k] = (P.avgproc == 1) ? 4. / (mu_ip1j + mu_ij + mu_ijp1 + mu_ip1jp1)
                                   : 1. / 4 * (1. / mu_ip1j + 1. / mu_ij + 1. / mu_ijp1 + 1. / mu_ip1jp1);
        return;
    }
    const double x = P.x_lb + ((i - 1) + 0.5) * P.dx, y = P.y_lb + ((j - 1) + 0.5) * P.dy;
    {
        const double Ka = K11_fn(P, x + P.dx, y), Kb = K11_fn(P, x, y), Kc = K11_fn(P, x, y + P.dy), Kd = K11_fn(P, x + P.dx, y + P.dy);
        KV11[k] = (P.avgproc == 1) ? 4. / (mu_ip1j / Ka + mu_ij / Kb + mu_ijp1 / Kc + mu_ip1jp1 / Kd)
                                   : 1. / 4 * (Ka / mu_ip1j + Kb / mu_ij + Kc / mu_ijp1 + Kd / mu_ip1jp1);
    }
    {
        const double Ka = K22_fn(P, x + P.dx, y), Kb = K22_fn(P, x, y), Kc = K22_fn(P, x, y + P.dy), Kd = K22_fn(P, x + P.dx, y + P.dy);
        KV22[k] = (P.avgproc == 1) ? 4. / (mu_ip1j / Ka + mu_ij / Kb + mu_ijp1 / Kc + mu_ip1jp1 / Kd)
                                   : 1. / 4 * (Ka / mu_ip1j + Kb / mu_ij + Kc / mu_ijp1 + Kd / mu_ip1jp1);
    }
}

IMPEC_D double vM11(const DevParams &P, double K11, double K22) { return P.secamt2 * (K11 * P.cosa2 + K22 * P.sina2); }
IMPEC_D double vM12(const DevParams &P, double K11, double K22) { return P.secamt2 * (K22 * P.costsina - K11 * P.cosasint); }
IMPEC_D double vM22(const DevParams &P, double K11, double K22) { return P.secamt2 * (K11 * P.sint2 + K22 * P.cost2); }

// Links AW, AS, AV on i in [0,N1], j in [0,N2] from MU and the vertex mobilities: a5-a8.
//   parallel part  src/MGSolver.cpp:760-773, rotated part :777-841, blend :748-756, folding :305-515.
template <bool TEST>
__global__ void __launch_bounds__(256) k_links(Grid g, const double *__restrict__ MU, const double *__restrict__ KV11, const double *__restrict__ KV22,
                                               double *__restrict__ AW, double *__restrict__ AS, double *__restrict__ AV, DevParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > g.N1 || j > g.N2) return;
    const size_t k = gidx(g, i, j);
    const double nu = P.nu;
    const double inv_xi2 = P.dxi * P.dxi, inv_xe = 1. / (P.deta * P.dxi);
    const bool rot = (nu != 1.);
#define AVRAW(kk) ((1 - nu) * (vM11(P, KV11[kk], KV22[kk]) / inv_xi2))
#define M12AT(kk) (vM12(P, KV11[kk], KV22[kk]))
    // vertex link
    const bool at_lo = (i == 0 && g.wall_lo), at_hi = (i == g.N1 && g.wall_hi);      // on a physical x-wall?
    const bool vb = (at_lo || at_hi || j == 0 || j == g.N2);
    AV[k] = (rot && !vb) ? AVRAW(k) : 0.;
    // x-face link between (i-1,j) and (i,j)
    if (j < g.N2) {
        double w = 0.;
        if (!at_lo && !at_hi) {
            double Ka = 1., Kb = 1.;
            if (TEST) { const double x = P.x_lb + (i + 0.5) * P.dx, y = P.y_lb + (j + 0.5) * P.dy; Ka = K11_fn(P, x - P.dx, y); Kb = K11_fn(P, x, y); }
            const double Wp = 1. / (P.dx * P.dx) * face_mob<TEST>(MU[k - g.ld], MU[k], Ka, Kb, P.avgproc);
            double Wd = 0.;
            if (rot) Wd = -inv_xe * (M12AT(k + 1) + M12AT(k));               // M21(i-1/2,j+1/2) + M12(i-1/2,j-1/2)
            w = nu * Wp + (1 - nu) * Wd;
            if (rot && j == 0) w = w + AVRAW(k);                              // bottom row: W += SW
            if (rot && j == g.N2 - 1) w = w + AVRAW(k + 1);                   // top row:    W += NW
        }
        AW[k] = w;
    }
    // y-face link between (i,j-1) and (i,j)
    if (i < g.N1) {
        double s = 0.;
        if (j > 0 && j < g.N2) {
            double Ka = 1., Kb = 1.;
            if (TEST) { const double x = P.x_lb + (i + 0.5) * P.dx, y = P.y_lb + (j + 0.5) * P.dy; Ka = K22_fn(P, x, y - P.dy); Kb = K22_fn(P, x, y); }
            const double Sp = 1. / (P.dy * P.dy) * face_mob<TEST>(MU[k - 1], MU[k], Ka, Kb, P.avgproc);
            double Sd = 0.;
            if (rot) Sd = inv_xe * (M12AT(k) + M12AT(k + g.ld));              // M12(i-1/2,j-1/2) + M21(i+1/2,j-1/2)
            s = nu * Sp + (1 - nu) * Sd;
            if (rot && i == 0 && g.wall_lo) s = s + AVRAW(k);                 // left column:  S += SW
            if (rot && i == g.N1 - 1 && g.wall_hi) s = s + AVRAW(k + g.ld);   // right column: S += SE
        }
        AS[k] = s;
    }
#undef AVRAW
#undef M12AT
}

// ---------------------------------------------------------------- multigrid level kernels
struct LevelDev {
    Grid g;
    const double *AW, *AS, *AV;
    const double *DINV;          // 1/diag (0 at the pinned cell), built once per operator by k_diag_inv
    // FP32 copies of the four coefficient arrays, read by the multigrid cycle kernels only (the arithmetic stays FP64
    // and the PCG mat-vec uses the FP64 links, so this changes the preconditioner by ~6e-8 relative, not the solution)
    const float *AWf, *ASf, *AVf, *DINVf;
    double mass;                 // 0 elliptic, 1 Crank-Nicolson
    // Dirichlet (pinned) cells: a contiguous index range [pin, pin_end) of the padded layout -- the single cell (1,1) of the parallel /
    // diagonal configurations, or a whole row (source_type horizontal: row i = 1 of the pressure problem, row N1-2 of the Crank-Nicolson
    // problem; src/MGSolver.cpp:109-140).  pin = pin_end = -1: none.
    long long pin, pin_end;
    // coarse levels: the cells [sig_idx, sig_end) carry the pins' diagonal shift sig[k - sig_idx] (k_pin_sigma / k_pin_sigma_row)
    long long sig_idx, sig_end;
    const double *sig;           // device pointer to the shifts: kernel arguments stay constant across steps
    // Row-chunk map of the marching kernels: CTA row blockIdx.y works on chunk cy0 + blockIdx.y * cys.  (0, 1) = all chunks in order; a
    // slab level launches the first and the last chunk -- the only ones that read halo rows -- as (0, ny - 1) x 2 behind the halo
    // exchange on a second stream, and the interior chunks (1, 1) x (ny - 2) concurrently with that exchange.
    int cy0, cys;
    // Marching kernels: software prefetch into L2 of the rows `pfd` steps ahead of the march (0 = off).  The register prefetch queue is
    // only 2 rows deep (register pressure); an L2 prefetch costs no register and turns the DRAM latency of those loads into L2 latency.
    int pfd;
};

IMPEC_D int chunk_of(const LevelDev &L) { return L.cy0 + (int)blockIdx.y * L.cys; }
IMPEC_D bool is_pin(const LevelDev &L, long long k) { return k >= L.pin && k < L.pin_end; }
IMPEC_D bool has_sig(const LevelDev &L, long long k) { return k >= L.sig_idx && k < L.sig_end; }

IMPEC_D double link_diag(const LevelDev &L, size_t k) {
    const int ld = L.g.ld;
    double d = L.mass + (((L.AW[k] + L.AW[k + ld]) + (L.AS[k] + L.AS[k + 1])) + ((L.AV[k] + L.AV[k + ld]) + (L.AV[k + 1] + L.AV[k + ld + 1])));
    if (has_sig(L, (long long)k)) d += L.sig[(long long)k - L.sig_idx];
    return d;
}
// order W,E,S,N,SW,SE,NW,NE (same as the oracle's off_sum)
IMPEC_D double link_offsum(const LevelDev &L, const double *__restrict__ x, size_t k) {
    const int ld = L.g.ld;
    double s = L.AW[k] * x[k - ld];
    s = fma(L.AW[k + ld], x[k + ld], s);
    s = fma(L.AS[k], x[k - 1], s);
    s = fma(L.AS[k + 1], x[k + 1], s);
    s = fma(L.AV[k], x[k - ld - 1], s);
    s = fma(L.AV[k + ld], x[k + ld - 1], s);
    s = fma(L.AV[k + 1], x[k - ld + 1], s);
    s = fma(L.AV[k + ld + 1], x[k + ld + 1], s);
    return s;
}

// One colour pass of red-black Gauss-Seidel (Jacobi inside a colour: the same-colour diagonal neighbours
// of the 9-point operator are read from xin).  Reads xin, writes xout for every cell (other colour copied),
// so a sweep = two passes ping-ponging x -> t -> x.
template <bool ZERO>
__global__ void __launch_bounds__(256) k_smooth_color(LevelDev L, int color, const double *__restrict__ f, const double *__restrict__ xin, double *__restrict__ xout) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= L.g.N1 || j >= L.g.N2) return;
    const size_t k = gidx(L.g, i, j);
    if (((i + j) & 1) != color || is_pin(L, (long long)k)) { xout[k] = ZERO ? 0. : xin[k]; return; }
    xout[k] = ZERO ? f[k] * L.DINV[k] : (f[k] + link_offsum(L, xin, k)) * L.DINV[k];
}

// FP32 copies of the coefficient arrays over ALL allocated rows (halo rows included, so no extra exchange)
__global__ void __launch_bounds__(256) k_links_to_f32(long long n, const double *__restrict__ AW, const double *__restrict__ AS, const double *__restrict__ AV,
                                                      const double *__restrict__ DINV, float *__restrict__ AWf, float *__restrict__ ASf, float *__restrict__ AVf, float *__restrict__ DINVf) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    AWf[k] = (float)AW[k]; ASf[k] = (float)AS[k]; AVf[k] = (float)AV[k]; DINVf[k] = (float)DINV[k];
}

// DINV = 1/diag on the interior (0 at the pinned cell); sigma (the pin's diagonal shift on coarse levels) is read
// from device memory so that no host round trip is needed between k_pin_sigma and this kernel.
__global__ void __launch_bounds__(256) k_diag_inv(LevelDev L, const double *__restrict__ sigma, double *__restrict__ DINV) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= L.g.N1 || j >= L.g.N2) return;
    const size_t k = gidx(L.g, i, j);
    const int ld = L.g.ld;
    double d = L.mass + (((L.AW[k] + L.AW[k + ld]) + (L.AS[k] + L.AS[k + 1])) + ((L.AV[k] + L.AV[k + ld]) + (L.AV[k + 1] + L.AV[k + ld + 1])));
    if (has_sig(L, (long long)k) && sigma) d += sigma[(long long)k - L.sig_idx];
    DINV[k] = is_pin(L, (long long)k) ? 0. : 1. / d;
}

// r = f - S x  (0 at the pinned cell); optionally accumulates <r,r> partials
__global__ void __launch_bounds__(256) k_residual(LevelDev L, const double *__restrict__ f, const double *__restrict__ x, double *__restrict__ r) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(blockIdx.y * blockDim.y + threadIdx.y) - (L.g.wall_lo ? 0 : 1);     // one halo row where a slab continues
    if (i >= L.g.N1 + (L.g.wall_hi ? 0 : 1) || j >= L.g.N2) return;
    const size_t k = gidx(L.g, i, j);
    r[k] = is_pin(L, (long long)k) ? 0. : f[k] - (link_diag(L, k) * x[k] - link_offsum(L, x, k));
}

// Residual r = f - S x when x is zero except at the pinned cell: r equals f everywhere but at the pin (0) and its 8 neighbours, so
// the host copies f and this kernel (9 threads) evaluates k_residual's expression on the 3 x 3 block around the pin.
__global__ void k_residual_pin_fix(LevelDev L, const double *__restrict__ f, const double *__restrict__ x, double *__restrict__ r) {
    const int t = threadIdx.x;
    if (t >= 9 || L.pin < 0) return;
    const int pi = (int)(L.pin / L.g.ld) - IMPEC_GH, pj = (int)(L.pin % L.g.ld) - IMPEC_PJ;
    const int i = pi + t / 3 - 1, j = pj + t % 3 - 1;
    if (i < 0 || i >= L.g.N1 || j < 0 || j >= L.g.N2) return;
    const size_t k = gidx(L.g, i, j);
    r[k] = is_pin(L, (long long)k) ? 0. : f[k] - (link_diag(L, k) * x[k] - link_offsum(L, x, k));
}

// || f - S x ||^2 without storing the residual (the true residual reported with every solve): 40 B/cell in one pass instead of
// k_residual + k_dot (48 + 16).  A CTA walks `rows` rows of a 32-column strip, so there are few partial sums.
__global__ void __launch_bounds__(256) k_residual_norm(LevelDev L, const double *__restrict__ f, const double *__restrict__ x, double *__restrict__ part, int rows) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.;
    for (int r = 0; r < rows; r += (int)blockDim.y) {
        const int i = (int)blockIdx.y * rows + r + (int)threadIdx.y;
        if (i < L.g.N1 && j < L.g.N2) {
            const size_t k = gidx(L.g, i, j);
            if (!is_pin(L, (long long)k)) { const double rr = f[k] - (link_diag(L, k) * x[k] - link_offsum(L, x, k)); v += rr * rr; }
        }
    }
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// full weighting = (bilinear prolongation)^T / 4 with mirrored ghosts (Neumann consistent)
__global__ void __launch_bounds__(256) k_restrict(Grid gf, const double *__restrict__ r, Grid gc, double *__restrict__ fc) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= gc.N1 || J >= gc.N2) return;
    const double w[4] = {0.25, 0.75, 0.75, 0.25};
    double t4[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        int i = 2 * I - 1 + a;
        if (i < 0 && gf.wall_lo) i = 0;
        if (i > gf.N1 - 1 && gf.wall_hi) i = gf.N1 - 1;
        double s = 0.;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int j = 2 * J - 1 + b; j = j < 0 ? 0 : (j > gf.N2 - 1 ? gf.N2 - 1 : j);
            s += w[b] * r[gidx(gf, i, j)];
        }
        t4[a] = s;
    }
    fc[gidx(gc, I, J)] = 0.25 * (w[0] * t4[0] + w[1] * t4[1] + w[2] * t4[2] + w[3] * t4[3]);
}

// x += P e : cell-centred bilinear interpolation with edge replication
__global__ void __launch_bounds__(256) k_prolong_add(Grid gc, const double *__restrict__ e, Grid gf, double *__restrict__ x, long long pin, long long pin_end) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= gf.N1 || j >= gf.N2) return;
    const size_t k = gidx(gf, i, j);
    if ((long long)k >= pin && (long long)k < pin_end) return;
    const int I = i >> 1, J = j >> 1;
    int In = (i & 1) ? I + 1 : I - 1;
    if (In < 0 && gc.wall_lo) In = 0;
    if (In > gc.N1 - 1 && gc.wall_hi) In = gc.N1 - 1;
    int Jn = (j & 1) ? J + 1 : J - 1; Jn = Jn < 0 ? 0 : (Jn > gc.N2 - 1 ? gc.N2 - 1 : Jn);
    const double a = 0.75 * e[gidx(gc, I, J)] + 0.25 * e[gidx(gc, In, J)];
    const double b = 0.75 * e[gidx(gc, I, Jn)] + 0.25 * e[gidx(gc, In, Jn)];
    x[k] += 0.75 * a + 0.25 * b;
}

// coarse links by rediscretisation on averaged fine links ("coarse operators")
__global__ void __launch_bounds__(256) k_coarsen_links(Grid gf, const double *__restrict__ AWf, const double *__restrict__ ASf, const double *__restrict__ AVf,
                                                       Grid gc, double *__restrict__ AWc, double *__restrict__ ASc, double *__restrict__ AVc) {
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y * blockDim.y + threadIdx.y;
    if (I > gc.N1 || J > gc.N2) return;
    const size_t kc = gidx(gc, I, J), kf = gidx(gf, 2 * I, 2 * J);
    const bool wall = (I == gc.N1 && gc.wall_hi);
    if (J < gc.N2) AWc[kc] = wall ? 0. : 0.125 * (AWf[kf] + AWf[kf + 1]);
    if (I < gc.N1) ASc[kc] = (J == gc.N2) ? 0. : 0.125 * (ASf[kf] + ASf[kf + gf.ld]);
    AVc[kc] = (wall || J == gc.N2) ? 0. : 0.25 * AVf[kf];
}

// ---------------------------------------------------------------- PCG vector kernels (a10)
// s = S d and partial <s,d>
__global__ void __launch_bounds__(256) k_matvec_dot(LevelDev L, const double *__restrict__ d, double *__restrict__ s, double *__restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double v = 0.;
    if (i < L.g.N1 && j < L.g.N2) {
        const size_t k = gidx(L.g, i, j);
        const double sk = is_pin(L, (long long)k) ? 0. : link_diag(L, k) * d[k] - link_offsum(L, d, k);
        s[k] = sk;
        v = sk * d[k];
    }
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// x += alpha d ; r -= alpha s ; partial <r,r>.  alpha = scal[0]/scal[1] is read from device memory.
// A pure stream (48 B/cell): a thread owns a column PAIR in UX_ROWS rows and issues all its 128-bit loads before the
// first store (256 B in flight per thread), and a CTA covers 256 x UX_ROWS cells, so there are 4x fewer partial sums
// for k_reduce_final than with one cell per thread.
#define UX_ROWS 4
#define UX_BLOCK 128
// MODE 0: x += alpha d, r -= alpha s (48 B/cell).
// The solution x is read by nobody until the solve ends, so its update can be DEFERRED and two of them merged (the search direction
// ping-pongs between two buffers, so d_k is still intact when d_{k+1} exists):
// MODE 1: r -= alpha s only, alpha is kept in aprev[0] (24 B/cell);
// MODE 2: x = (x + aprev[0] dprev) + alpha d, r -= alpha s (56 B/cell) -- 80 instead of 96 B/cell per two iterations, and the same
//         operations on x in the same order as two MODE 0 passes: bit-identical.
template <int MODE>
__global__ void __launch_bounds__(UX_BLOCK) k_update_xr(Grid g, const double *__restrict__ scal, const double *__restrict__ d, const double *__restrict__ s,
                                                        double *__restrict__ x, double *__restrict__ r, double *__restrict__ part,
                                                        const double *__restrict__ dprev, double *__restrict__ aprev) {
    const int j = 2 * (int)(blockIdx.x * UX_BLOCK + threadIdx.x);
    const int i0 = (int)blockIdx.y * UX_ROWS;
    const double alpha = scal[0] / scal[1];
    const double alpha0 = (MODE == 2) ? aprev[0] : 0.;
    double v = 0.;
    if (j + 1 < g.N2) {
        d2 dv[UX_ROWS], sv[UX_ROWS], xv[UX_ROWS], rv[UX_ROWS], pv[UX_ROWS];
#pragma unroll
        for (int n = 0; n < UX_ROWS; n++) {
            if (i0 + n >= g.N1) continue;
            const size_t k = gidx(g, i0 + n, j);
            sv[n] = ld2(s + k); rv[n] = ld2(r + k);
            if (MODE != 1) { dv[n] = ld2(d + k); xv[n] = ld2(x + k); }
            if (MODE == 2) pv[n] = ld2(dprev + k);
        }
#pragma unroll
        for (int n = 0; n < UX_ROWS; n++) {
            if (i0 + n >= g.N1) continue;
            const size_t k = gidx(g, i0 + n, j);
            if (MODE == 0) st2(x + k, xv[n].x + alpha * dv[n].x, xv[n].y + alpha * dv[n].y);
            if (MODE == 2) st2(x + k, (xv[n].x + alpha0 * pv[n].x) + alpha * dv[n].x, (xv[n].y + alpha0 * pv[n].y) + alpha * dv[n].y);
            const double ra = rv[n].x - alpha * sv[n].x, rb = rv[n].y - alpha * sv[n].y;
            st2(r + k, ra, rb);
            v += ra * ra; v += rb * rb;
        }
    } else if (j < g.N2) {
        for (int n = 0; n < UX_ROWS && i0 + n < g.N1; n++) {
            const size_t k = gidx(g, i0 + n, j);
            if (MODE == 0) x[k] += alpha * d[k];
            if (MODE == 2) x[k] = (x[k] + alpha0 * dprev[k]) + alpha * d[k];
            const double rk = r[k] - alpha * s[k];
            r[k] = rk;
            v += rk * rk;
        }
    }
    if (MODE == 1 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) aprev[0] = alpha;
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}
// x += a[0] d   (flush of a deferred solution update)
__global__ void __launch_bounds__(UX_BLOCK) k_flush_x(Grid g, const double *__restrict__ a, const double *__restrict__ d, double *__restrict__ x) {
    const int j = 2 * (int)(blockIdx.x * UX_BLOCK + threadIdx.x);
    const int i0 = (int)blockIdx.y * UX_ROWS;
    const double alpha = a[0];
    for (int n = 0; n < UX_ROWS && i0 + n < g.N1; n++) {
        const size_t k = gidx(g, i0 + n, j);
        if (j + 1 < g.N2) { const d2 dv = ld2(d + k), xv = ld2(x + k); st2(x + k, xv.x + alpha * dv.x, xv.y + alpha * dv.y); }
        else if (j < g.N2) x[k] += alpha * d[k];
    }
}

// partial <a,b> (and optionally <a,a> in a second partial array)
__global__ void __launch_bounds__(256) k_dot(Grid g, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double v = 0.;
    if (i < g.N1 && j < g.N2) { const size_t k = gidx(g, i, j); v = a[k] * b[k]; }
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// d = z + beta d, beta = scal[0]/scal[1]   (FIRST: d = z)
template <bool FIRST>
__global__ void __launch_bounds__(256) k_update_d(Grid g, const double *__restrict__ scal, const double *__restrict__ z, double *__restrict__ d) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.N1 || j >= g.N2) return;
    const size_t k = gidx(g, i, j);
    d[k] = FIRST ? z[k] : z[k] + (scal[0] / scal[1]) * d[k];
}

// b = -q scattered into the padded layout, x0 = 0 except the pinned value; partial <b,b>
__global__ void __launch_bounds__(256) k_setup_rhs(Grid g, const double *__restrict__ q, double *__restrict__ rhs, double *__restrict__ x,
                                                   long long pin, long long pin_end, double pin_value, double pin_scale, double *__restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double v = 0.;
    if (i < g.N1 && j < g.N2) {
        const size_t k = gidx(g, i, j);
        const double qk = q[k];
        rhs[k] = qk;                       // SPD form: S x = -b = q
        // Dirichlet cells: one cell with a given value, or a row whose values follow from the row's own equation kappa/dA p = -q (pin_scale = -dA)
        x[k] = ((long long)k >= pin && (long long)k < pin_end) ? (pin_scale != 0. ? pin_scale * qk : pin_value) : 0.;
        v = qk * qk;                       // <b,b> = <q,q>
    }
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// ---------------------------------------------------------------- a19: Crank-Nicolson dispersion
// face dispersion coefficients (src/MGSolver.cpp:846-919) turned into links of (I - dt/2 L_D):
//   AW(i,j) = dt/(2 dx^2) d11(i-1/2,j),  AS(i,j) = dt/(2 dy^2) d22(i,j-1/2),  AV = 0   (alpha_l == alpha_t => d12 = d21 = 0)
// |u| = 0 gives 0/0 upstream (:868-869); here such a face simply has no dispersion.
__global__ void __launch_bounds__(256) k_disp_links(Grid g, const double *__restrict__ u, const double *__restrict__ v, double dt, DevParams P,
                                                    double *__restrict__ AW, double *__restrict__ AS, double *__restrict__ AV) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > g.N1 || j > g.N2) return;
    const size_t k = gidx(g, i, j);
    const int ld = g.ld;
    const double maxdx = P.dx > P.dy ? P.dx : P.dy;
    const double al = P.alpha_l * maxdx, at = P.alpha_t * maxdx;
    AV[k] = 0.;
    if (j < g.N2) {
        double w = 0.;
        if (!(i == 0 && g.wall_lo) && !(i == g.N1 && g.wall_hi)) {
            const double uu = u[k];
            const double vv = 1. / 4 * (v[k - ld + 1] + v[k + 1] + v[k - ld] + v[k]);
            const double u2 = uu * uu, v2 = vv * vv, n2 = u2 + v2;
            if (n2 > 0.) {
                const double fl = P.porosity * al / sqrt(n2), ft = P.porosity * at / sqrt(n2);
                w = dt / (2. * P.dx * P.dx) * (fl * u2 + ft * v2);
            }
        }
        AW[k] = w;
    }
    if (i < g.N1) {
        double s = 0.;
        if (j > 0 && j < g.N2) {
            const double uu = 1. / 4 * (u[k + ld - 1] + u[k + ld] + u[k - 1] + u[k]);
            const double vv = v[k];
            const double u2 = uu * uu, v2 = vv * vv, n2 = u2 + v2;
            if (n2 > 0.) {
                const double fl = P.porosity * al / sqrt(n2), ft = P.porosity * at / sqrt(n2);
                s = dt / (2. * P.dy * P.dy) * (fl * v2 + ft * u2);
            }
        }
        AS[k] = s;
    }
}
// rhs = (I + dt/2 L_D) c = c + sum links (c_nb - c)   (src/MGSolver.cpp:1324-1371), x0 = 0, partial <b,b>
__global__ void __launch_bounds__(256) k_cn_rhs(LevelDev L, const double *__restrict__ c, double *__restrict__ rhs, double *__restrict__ x, double *__restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double v = 0.;
    if (i < L.g.N1 && j < L.g.N2) {
        const size_t k = gidx(L.g, i, j);
        const double links = link_diag(L, k) - L.mass;
        const double b = is_pin(L, (long long)k) ? 0. : c[k] + (link_offsum(L, c, k) - links * c[k]);      // Dirichlet rows: rhs = 0 (src/MGSolver.cpp:1357-1364)
        rhs[k] = b; x[k] = 0.;
        v = b * b;
    }
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// ---- general dispersion tensor (alpha_l != alpha_t): the Crank-Nicolson matrix has all nine entries and is NOT symmetric
// (src/MGSolver.cpp:923-948).  Stored as the four face arrays it is built from -- d11, d12 on x faces (SetupDiffusionTensorF, :846-882),
// d22, d21 on y faces (SetupDiffusionTensorG, :886-919) -- and applied matrix-free.  The Neumann folding of :1031-1289 adds every entry
// whose neighbour lies outside to the entry of that neighbour's mirror image, i.e. A_folded x = A_unfolded x~ with x~ the field with
// mirrored ghost cells: the operator kernel below reads a ghost-filled x.
__global__ void __launch_bounds__(256) k_cn9_faces(Grid g, const double *__restrict__ u, const double *__restrict__ v, DevParams P,
                                                   double *__restrict__ F11, double *__restrict__ F12, double *__restrict__ G22, double *__restrict__ G21) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > g.N1 || j > g.N2) return;
    const size_t k = gidx(g, i, j);
    const int ld = g.ld;
    const double maxdx = P.dx > P.dy ? P.dx : P.dy;
    const double al = P.alpha_l * maxdx, at = P.alpha_t * maxdx;
    if (j < g.N2) {
        const double uu = u[k];
        const double vv = 1. / 4 * (v[k - ld + 1] + v[k + 1] + v[k - ld] + v[k]);
        const double u2 = uu * uu, v2 = vv * vv, n2 = u2 + v2;
        double d11 = 0., d12 = 0.;
        if (n2 > 0.) {                                   // |u| = 0 gives 0/0 upstream (:868-869); no dispersion here
            const double fl = P.porosity * al / sqrt(n2), ft = P.porosity * at / sqrt(n2);
            d11 = fl * u2 + ft * v2;
            d12 = fl * (uu * vv) + ft * (-uu * vv);
        }
        F11[k] = d11; F12[k] = d12;
    }
    if (i < g.N1) {
        const double uu = 1. / 4 * (u[k + ld - 1] + u[k + ld] + u[k - 1] + u[k]);
        const double vv = v[k];
        const double u2 = uu * uu, v2 = vv * vv, n2 = u2 + v2;
        double d22 = 0., d21 = 0.;
        if (n2 > 0.) {
            const double fl = P.porosity * al / sqrt(n2), ft = P.porosity * at / sqrt(n2);
            d21 = fl * (uu * vv) + ft * (-uu * vv);
            d22 = fl * v2 + ft * u2;
        }
        G22[k] = d22; G21[k] = d21;
    }
}
struct Cn9 { const double *F11, *F12, *G22, *G21; double dt; };
// the nine entries of row (i,j) in the order C, W, SW, S, SE, E, NE, N, NW (src/MGSolver.cpp:286-294, 923-948)
IMPEC_D void cn9_row(const Grid &g, const Cn9 &A, const DevParams &P, size_t k, double e[9]) {
    const int ld = g.ld;
    const double dx = P.dx, dy = P.dy, dt = A.dt;
    const double d11_im = A.F11[k], d12_im = A.F12[k], d11_ip = A.F11[k + ld], d12_ip = A.F12[k + ld];
    const double d22_jm = A.G22[k], d21_jm = A.G21[k], d22_jp = A.G22[k + 1], d21_jp = A.G21[k + 1];
    e[8] = dt / (8. * dx * dy) * (d12_im + d21_jp);                                                    // NW
    e[6] = -dt / (8. * dx * dy) * (d12_ip + d21_jp);                                                   // NE
    e[7] = -dt / 8. * (4. / (dy * dy) * d22_jp + 1. / (dx * dy) * (-d12_im + d12_ip));                 // N
    e[2] = -dt / (8. * dx * dy) * (d12_im + d21_jm);                                                   // SW
    e[4] = dt / (8. * dx * dy) * (d12_ip + d21_jm);                                                    // SE
    e[3] = -dt / 8. * (4. / (dy * dy) * d22_jm + 1. / (dx * dy) * (-d12_ip + d12_im));                 // S
    e[1] = -dt / 8. * (4. / (dx * dx) * d11_im + 1. / (dx * dy) * (-d21_jp + d21_jm));                 // W
    e[0] = 1. + dt / (2. * dx * dx) * (d11_ip + d11_im) + dt / (2. * dy * dy) * (d22_jp + d22_jm);     // C
    e[5] = -dt / 8. * (4. / (dx * dx) * d11_ip + 1. / (dx * dy) * (-d21_jm + d21_jp));                 // E
}
// y = A x (x with mirrored ghost cells).  DIAG != nullptr: also the inverse of the FOLDED diagonal (the Jacobi preconditioner)
__global__ void __launch_bounds__(256) k_cn9_apply(Grid g, Cn9 A, DevParams P, const double *__restrict__ x, double *__restrict__ y, double *__restrict__ dinv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.N1 || j >= g.N2) return;
    const size_t k = gidx(g, i, j);
    const int ld = g.ld;
    double e[9];
    cn9_row(g, A, P, k, e);
    const int di[9] = {0, -1, -1, 0, 1, 1, 1, 0, -1}, dj[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1};
    double s = 0., dg = 0.;
#pragma unroll
    for (int a = 0; a < 9; a++) {
        s += e[a] * x[(long long)k + (long long)di[a] * ld + dj[a]];
        int ii = i + di[a], jj = j + dj[a];
        if (ii < 0 && g.wall_lo) ii = 0;                       // (a slab edge that is no wall folds nothing: the neighbour row is a real row)
        if (ii > g.N1 - 1 && g.wall_hi) ii = g.N1 - 1;
        jj = jj < 0 ? 0 : (jj > g.N2 - 1 ? g.N2 - 1 : jj);
        if (ii == i && jj == j) dg += e[a];
    }
    y[k] = s;
    if (dinv) dinv[k] = 1. / dg;
}
// b = (I + dt/2 L_D) c with the full tensor (src/MGSolver.cpp:1324-1355), x0 = 0, partial <b,b>; c carries filled ghosts
__global__ void __launch_bounds__(256) k_cn9_rhs(Grid g, Cn9 A, DevParams P, const double *__restrict__ c, double *__restrict__ rhs, double *__restrict__ x, double *__restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double v = 0.;
    if (i < g.N1 && j < g.N2) {
        const size_t k = gidx(g, i, j);
        const int ld = g.ld;
        const double dx = P.dx, dy = P.dy;
        const double d11_im = A.F11[k], d12_im = A.F12[k], d11_ip = A.F11[k + ld], d12_ip = A.F12[k + ld];
        const double d22_jm = A.G22[k], d21_jm = A.G21[k], d22_jp = A.G22[k + 1], d21_jp = A.G21[k + 1];
#define C_(a, b) c[(long long)k + (long long)(a) * ld + (b)]
        const double b = C_(0, 0) + A.dt / 8. * (
            4. / (dx * dx) * (d11_ip * (C_(1, 0) - C_(0, 0)) - d11_im * (C_(0, 0) - C_(-1, 0)))
            + 1. / (dx * dy) * (d12_ip * (C_(0, 1) + C_(1, 1) - C_(0, -1) - C_(1, -1)) - d12_im * (C_(0, 1) + C_(-1, 1) - C_(-1, -1) - C_(0, -1)))
            + 1. / (dx * dy) * (d21_jp * (C_(1, 0) + C_(1, 1) - C_(-1, 1) - C_(-1, 0)) - d21_jm * (C_(1, 0) + C_(1, -1) - C_(-1, 0) - C_(-1, -1)))
            + 4. / (dy * dy) * (d22_jp * (C_(0, 1) - C_(0, 0)) - d22_jm * (C_(0, 0) - C_(0, -1))));
#undef C_
        rhs[k] = b; x[k] = 0.;
        v = b * b;
    }
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}
// BiCGStab vector updates (interior cells)
__global__ void __launch_bounds__(256) k_bicg_p(Grid g, double beta, double omega, const double *__restrict__ r, const double *__restrict__ v, const double *__restrict__ dinv,
                                                double *__restrict__ p, double *__restrict__ ph) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.N1 || j >= g.N2) return;
    const size_t k = gidx(g, i, j);
    const double pn = r[k] + beta * (p[k] - omega * v[k]);
    p[k] = pn; ph[k] = pn * dinv[k];
}
__global__ void __launch_bounds__(256) k_bicg_s(Grid g, double alpha, const double *__restrict__ r, const double *__restrict__ v, const double *__restrict__ dinv,
                                                double *__restrict__ s, double *__restrict__ sh) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.N1 || j >= g.N2) return;
    const size_t k = gidx(g, i, j);
    const double sn = r[k] - alpha * v[k];
    s[k] = sn; sh[k] = sn * dinv[k];
}
__global__ void __launch_bounds__(256) k_bicg_xr(Grid g, double alpha, double omega, const double *__restrict__ ph, const double *__restrict__ sh, const double *__restrict__ s,
                                                 const double *__restrict__ t, double *__restrict__ x, double *__restrict__ r, double *__restrict__ part) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double v = 0.;
    if (i < g.N1 && j < g.N2) {
        const size_t k = gidx(g, i, j);
        x[k] += alpha * ph[k] + omega * sh[k];
        const double rn = s[k] - omega * t[k];
        r[k] = rn; v = rn * rn;
    }
    v = block_reduce<RED_SUM>(v);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = v;
}

// ---------------------------------------------------------------- a12: Darcy velocity
// rotated-frame velocity at vertex (i-1/2,j-1/2) rotated back to (x,y): src/EllipticSolver.cpp:554-592
IMPEC_D void vertex_vel(const DevParams &P, const double *__restrict__ p, double K11, double K22, size_t k, int ld, double &u1, double &u2) {
    const double d1dp = 1. / P.dxi * (p[k] - p[k - ld - 1]);
    const double d2dp = 1. / P.deta * (p[k - ld] - p[k - 1]);
    const double M11 = vM11(P, K11, K22), M12 = vM12(P, K11, K22), M22 = vM22(P, K11, K22);
    const double w1 = -d1dp * M11 - d2dp * M12;
    const double w2 = -d1dp * M12 - d2dp * M22;
    u1 = P.cost * w1 - P.sina * w2;
    u2 = P.sint * w1 + P.cosa * w2;
}

// u on i in [0,N1], j in [0,N2); v on i in [0,N1), j in [0,N2]: src/EllipticSolver.cpp:541-547, 600-683.
// p must carry one filled Neumann ghost ring.
template <bool TEST>
__global__ void __launch_bounds__(256) k_velocity(Grid g, const double *__restrict__ p, const double *__restrict__ MU, const double *__restrict__ KV11,
                                                  const double *__restrict__ KV22, double *__restrict__ u, double *__restrict__ v, DevParams P) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > g.N1 || j > g.N2) return;
    const size_t k = gidx(g, i, j);
    const int ld = g.ld;
    const double nu = P.nu;
    const bool rot = (nu != 1.);
    double a1 = 0., a2 = 0.;
    if (rot) vertex_vel(P, p, KV11[k], KV22[k], k, ld, a1, a2);
    if (j < g.N2) {
        double Ka = 1., Kb = 1.;
        if (TEST) { const double x = P.x_lb + (i + 0.5) * P.dx, y = P.y_lb + (j + 0.5) * P.dy; Ka = K11_fn(P, x - P.dx, y); Kb = K11_fn(P, x, y); }
        const double par = -face_mob<TEST>(MU[k - ld], MU[k], Ka, Kb, P.avgproc) * (p[k] - p[k - ld]) / P.dx;
        double b1 = 0., b2 = 0.;
        if (rot) vertex_vel(P, p, KV11[k + 1], KV22[k + 1], k + 1, ld, b1, b2);
        u[k] = nu * par + (1. - nu) * (0.5 * (b1 + a1));
    }
    if (i < g.N1) {
        double Ka = 1., Kb = 1.;
        if (TEST) { const double x = P.x_lb + (i + 0.5) * P.dx, y = P.y_lb + (j + 0.5) * P.dy; Ka = K22_fn(P, x, y - P.dy); Kb = K22_fn(P, x, y); }
        const double par = -face_mob<TEST>(MU[k - 1], MU[k], Ka, Kb, P.avgproc) * (p[k] - p[k - 1]) / P.dy;
        double b1 = 0., b2 = 0.;
        if (rot) vertex_vel(P, p, KV11[k + ld], KV22[k + ld], k + ld, ld, b1, b2);
        v[k] = nu * par + (1. - nu) * (0.5 * (a2 + b2));
    }
}

// ---------------------------------------------------------------- diagnostics (a20)
// partial sums: [0] sum c, [1] sum |div v - q|, plus boundary velocity handled by a small kernel
__global__ void __launch_bounds__(256) k_diag_cells(Grid g, const double *__restrict__ c, const double *__restrict__ u, const double *__restrict__ v,
                                                    const double *__restrict__ q, double dx, double dy, double *__restrict__ part, long long stride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double m = 0., e = 0., cmax = -DBL_MAX, cmin = DBL_MAX;
    if (i < g.N1 && j < g.N2) {
        const size_t k = gidx(g, i, j);
        m = c[k]; cmax = m; cmin = m;
        const double div = (v[k + 1] - v[k]) / dy + (u[k + g.ld] - u[k]) / dx;
        e = fabs(div - q[k]);
    }
    m = block_reduce<RED_SUM>(m);
    e = block_reduce<RED_SUM>(e);
    cmax = block_reduce<RED_MAX>(cmax);
    cmin = block_reduce<RED_MIN>(cmin);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const long long b = blockIdx.y * gridDim.x + blockIdx.x;
        part[b] = m; part[stride + b] = e; part[2 * stride + b] = cmax; part[3 * stride + b] = cmin;
    }
}
// Integral diagnostics of a displacement run (definitions: impec2d_b200/diagnostics.py) as device reductions:
// partial sums [0] number of cells with lo <= c <= hi (the mixing zone), [1] sum c.  One pass over c, no field leaves the GPU.
__global__ void __launch_bounds__(256) k_diag_mixing(Grid g, const double *__restrict__ c, double lo, double hi, double *__restrict__ part, long long stride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double n = 0., m = 0.;
    if (i < g.N1 && j < g.N2) {
        m = c[gidx(g, i, j)];
        n = (m >= lo && m <= hi) ? 1. : 0.;
    }
    n = block_reduce<RED_SUM>(n);
    m = block_reduce<RED_SUM>(m);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const long long b = blockIdx.y * gridDim.x + blockIdx.x;
        part[b] = n; part[stride + b] = m;
    }
}
// |q|-weighted concentration over the production cells (q < 0) of the compact well list: partial sums [0] sum -q, [1] sum -q c
__global__ void __launch_bounds__(256) k_diag_wells(int n, const int *__restrict__ wi, const int *__restrict__ wj, const double *__restrict__ wq, Grid g,
                                                    const double *__restrict__ c, double *__restrict__ part, long long stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double w = 0., wc = 0.;
    if (t < n) {
        const double q = wq[t];
        if (q < 0.) { w = -q; wc = w * c[gidx(g, wi[t], wj[t])]; }
    }
    w = block_reduce<RED_SUM>(w);
    wc = block_reduce<RED_SUM>(wc);
    if (threadIdx.x == 0) { part[blockIdx.x] = w; part[stride + blockIdx.x] = wc; }
}
// no-flow boundary velocity sum: src/Impec2DSolver.cpp:562-588
__global__ void __launch_bounds__(256) k_diag_boundary(Grid g, const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ part) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.;
    if (t < g.N1) s += fabs(v[gidx(g, t, 0)]) + fabs(v[gidx(g, t, g.N2)]);
    if (t < g.N2) s += (g.wall_lo ? fabs(u[gidx(g, 0, t)]) : 0.) + (g.wall_hi ? fabs(u[gidx(g, g.N1, t)]) : 0.);
    s = block_reduce<RED_SUM>(s);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
