/* impec_oracle.c -- CPU ORACLE (test infrastructure only; see impec_oracle.h).
 *
 * Restates, in plain C, what pw359/Impec2D computes on its IMPES hot path.
 * Citations are file:line under /root/reference/.  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp   (no FMA contraction: the CUDA kernels are
 *   built with -fmad=false so the explicit transport path can be compared bit for bit).
 */
#include "impec_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ indexing */
static inline int SP(const orc_params *P) { return P->N2 + 2 * P->g; }       /* scalar pitch   */
static inline int VP(const orc_params *P) { return P->N2 + 1 + 2 * P->g; }   /* velocity pitch */
static inline size_t ci(const orc_params *P, int i, int j) { return (size_t)(i + P->g) * SP(P) + (size_t)(j + P->g); }
static inline size_t vi(const orc_params *P, int i, int j) { return (size_t)(i + P->g) * VP(P) + (size_t)(j + P->g); }
static inline double DX(const orc_params *P) { return (P->x_ub - P->x_lb) / P->N1; }   /* SimParams.cpp:452-453 */
static inline double DY(const orc_params *P) { return (P->y_ub - P->y_lb) / P->N2; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void orc_default_params(orc_params *P) {
    memset(P, 0, sizeof(*P));
    P->N1 = P->N2 = 64; P->g = 2;
    P->x_lb = P->y_lb = 0.0; P->x_ub = P->y_ub = 1.0;
    P->mu_res = 41.0; P->mu_inj = 1.0; P->avgproc = ORC_HARMONIC;
    P->nu = 2.0 / 3.0;
    P->stype = ORC_PARALLEL;
    P->porosity = 1.0; P->cfl = 0.5; P->inj_fluid_conc = 1.0;
    P->alpha_l = P->alpha_t = 1.0; P->dt_diff_limit = 0.9; P->ptype = ORC_CRANKNICOLSON;
    P->eps = 1e-14; P->para_eps = 1e-14;
    P->hyper_method = ORC_MUSCL; P->dim_splitting = ORC_SYMMETRIC_PARALLEL;
    P->max_iter = 10000; P->mg_max_levels = 0; P->mg_nmin = 4; P->mg_ncoarse = 8;
    P->mg_w_from = 2; P->mg_w_to = 4;
}

void orc_state_init(orc_state *S) {
    memset(S, 0, sizeof(*S));
    S->prev_first_sweep = 1;            /* Y_SWEEP: HyperbolicSolver.cpp:84 */
}

/* ------------------------------------------------------------------ ghost fills */
/* src/Util.cpp:215-256 */
void orc_fill_periodic_scalar(double *a, int N1, int N2, int g) {
    const int pitch = N2 + 2 * g;
#define A_(i, j) a[(size_t)((i) + g) * pitch + ((j) + g)]
    for (int i = 0; i < N1; i++)
        for (int k = 0; k < g; k++) { A_(i, N2 + k) = A_(i, k); A_(i, -(k + 1)) = A_(i, N2 - 1 - k); }
    for (int k = 0; k < g; k++)
        for (int j = 0; j < N2; j++) { A_(-(k + 1), j) = A_(N1 - 1 - k, j); A_(N1 + k, j) = A_(k, j); }
    for (int k = 0; k < g; k++)
        for (int l = 0; l < g; l++) {
            A_(-(k + 1), -(l + 1)) = A_(N1 - 1 - k, N2 - 1 - l);
            A_(N1 + k, -(l + 1))   = A_(k, N2 - 1 - l);
            A_(-(k + 1), N2 + l)   = A_(N1 - 1 - k, l);
            A_(N1 + k, N2 + l)     = A_(k, l);
        }
#undef A_
}

/* src/Util.cpp:283-324 */
void orc_fill_neumann_scalar(double *a, int N1, int N2, int g) {
    const int pitch = N2 + 2 * g;
#define A_(i, j) a[(size_t)((i) + g) * pitch + ((j) + g)]
    for (int i = 0; i < N1; i++)
        for (int k = 0; k < g; k++) { A_(i, N2 + k) = A_(i, N2 - 1 - k); A_(i, -(k + 1)) = A_(i, k); }
    for (int k = 0; k < g; k++)
        for (int j = 0; j < N2; j++) { A_(-(k + 1), j) = A_(k, j); A_(N1 + k, j) = A_(N1 - 1 - k, j); }
    for (int k = 0; k < g; k++)
        for (int l = 0; l < g; l++) {
            A_(-(k + 1), -(l + 1)) = A_(k, l);
            A_(N1 + k, -(l + 1))   = A_(N1 - 1 - k, l);
            A_(-(k + 1), N2 + l)   = A_(k, N2 - 1 - l);
            A_(N1 + k, N2 + l)     = A_(N1 - 1 - k, N2 - 1 - l);
        }
#undef A_
}

/* src/Util.cpp:331-438.  u is antisymmetric about the walls i=0,N1 and mirrored in j;
 * v is antisymmetric about j=0,N2 and mirrored in i.  Statement order follows the reference
 * because later fills read earlier ones (corner strips). */
void orc_fill_neumann_velocity(double *u, double *v, int N1, int N2, int g) {
    const int pitch = N2 + 1 + 2 * g;
#define U_(i, j) u[(size_t)((i) + g) * pitch + ((j) + g)]
#define V_(i, j) v[(size_t)((i) + g) * pitch + ((j) + g)]
    for (int i = 0; i <= N1; i++)
        for (int k = 0; k < g; k++) { U_(i, -(k + 1)) = U_(i, k); U_(i, N2 + k) = U_(i, N2 - 1 - k); }
    for (int k = 0; k < g; k++)
        for (int l = 0; l < g; l++) { U_(-(k + 1), N2 + l) = -U_(k + 1, N2 + l); U_(-(k + 1), -(l + 1)) = -U_(k + 1, -(l + 1)); }
    for (int k = 0; k < g; k++)
        for (int l = 0; l < g; l++) { U_(N1 + k + 1, N2 + l) = -U_(N1 - 1 - k, N2 + l); U_(N1 + k + 1, -(l + 1)) = -U_(N1 - k - 1, -(l + 1)); }
    for (int k = 0; k < g; k++)
        for (int j = 0; j < N2; j++) { U_(-(k + 1), j) = -U_(k + 1, j); U_(N1 + k + 1, j) = -U_(N1 - k - 1, j); }

    for (int k = 0; k < g; k++)
        for (int j = 0; j <= N2; j++) { V_(-(k + 1), j) = V_(k, j); V_(N1 + k, j) = V_(N1 - k - 1, j); }
    for (int i = 0; i < N1; i++)
        for (int l = 0; l < g; l++) { V_(i, -(l + 1)) = -V_(i, l + 1); V_(i, N2 + l + 1) = -V_(i, N2 - 1 - l); }
    for (int k = 0; k < g; k++)
        for (int l = 0; l < g; l++) { V_(-(k + 1), N2 + 1 + l) = V_(k, N2 + 1 + l); V_(-(k + 1), -(l + 1)) = V_(k, -(l + 1)); }
    for (int k = 0; k < g; k++)
        for (int l = 0; l < g; l++) { V_(N1 + k, N2 + 1 + l) = V_(N1 - 1 - k, N2 + l + 1); V_(N1 + k, -(l + 1)) = V_(N1 - 1 - k, -(l + 1)); }
#undef U_
#undef V_
}

/* src/HyperbolicSolver.cpp:56-61 */
void orc_fill_boundaries(const orc_params *P, double *c) {
    if (P->is_test) orc_fill_periodic_scalar(c, P->N1, P->N2, P->g);
    else            orc_fill_neumann_scalar(c, P->N1, P->N2, P->g);
}

/* ------------------------------------------------------------------ mobility */
/* src/EllipticSolver.cpp:126, 692-695 */
double orc_mu(const orc_params *P, double c) {
    const double m = pow(P->mu_res / P->mu_inj, 0.25);
    const double x = (1. - c + m * c);
    return 1. / (x * x * x * x) * P->mu_res;
}

/* src/EllipticSolver.cpp:418-470 */
double orc_K11(const orc_params *P, double x, double y) {
    if (!P->is_test) return 1.;
    switch (P->testno) {
    case 1: return 1.;
    case 2: return exp(-0.3 * x);
    case 3: return 2. + sin(2. * x) * sin(y);
    case 4: return 2. + sin(2. * x) * sin(y);
    default: return NAN;
    }
}
double orc_K22(const orc_params *P, double x, double y) {
    if (!P->is_test) return 1.;
    switch (P->testno) {
    case 1: return 1.;
    case 2: return exp(-0.3 * y);
    case 3: return 2. + sin(2. * x) * sin(y);
    case 4: return 2. + sin(x) * sin(2. * y);
    default: return NAN;
    }
}

typedef struct trig_t {            /* src/EllipticSolver.cpp:59-93 */
    double dx, dy, dA, alpha, theta, dxi, deta, secamt2, cosa2, sina2, costsina, cosasint, cost2, sint2, cosa, sint, cost, sina;
} trig_t;

static trig_t make_trig(const orc_params *P) {
    trig_t t;
    t.dx = DX(P); t.dy = DY(P); t.dA = t.dx * t.dy;
    t.alpha = atan(t.dx / t.dy); t.theta = atan(t.dy / t.dx);
    t.dxi = t.dx / cos(t.theta); t.deta = t.dy / cos(t.alpha);
    const double secamt = 1. / cos(t.alpha - t.theta);
    t.secamt2 = secamt * secamt;
    t.cosa2 = cos(t.alpha) * cos(t.alpha); t.sina2 = sin(t.alpha) * sin(t.alpha);
    t.costsina = cos(t.theta) * sin(t.alpha); t.cosasint = cos(t.alpha) * sin(t.theta);
    t.cost2 = cos(t.theta) * cos(t.theta); t.sint2 = sin(t.theta) * sin(t.theta);
    t.cosa = cos(t.alpha); t.sint = sin(t.theta); t.cost = cos(t.theta); t.sina = sin(t.alpha);
    return t;
}

/* face mobility between cells (i-1,j) and (i,j): src/EllipticSolver.cpp:478-499 */
static double kx_face(const orc_params *P, double mq, double c_im1j, double c_ij, int i, int j) {
    const double dx = DX(P), dy = DY(P);
    const double x = P->x_lb + (i + 0.5) * dx, y = P->y_lb + (j + 0.5) * dy;
    const double Ka = orc_K11(P, x - dx, y), Kb = orc_K11(P, x, y);
    const double xa = (1. - c_im1j + mq * c_im1j), xb = (1. - c_ij + mq * c_ij);
    const double mu_b = 1. / (xb * xb * xb * xb) * P->mu_res, mu_a = 1. / (xa * xa * xa * xa) * P->mu_res;
    if (P->avgproc == ORC_HARMONIC) return 2. / (mu_a / Ka + mu_b / Kb);
    return 1. / 2 * (Ka / mu_a + Kb / mu_b);
}
/* face mobility between cells (i,j-1) and (i,j): src/EllipticSolver.cpp:501-524 */
static double ky_face(const orc_params *P, double mq, double c_ijm1, double c_ij, int i, int j) {
    const double dx = DX(P), dy = DY(P);
    const double x = P->x_lb + (i + 0.5) * dx, y = P->y_lb + (j + 0.5) * dy;
    const double Ka = orc_K22(P, x, y - dy), Kb = orc_K22(P, x, y);
    const double xa = (1. - c_ijm1 + mq * c_ijm1), xb = (1. - c_ij + mq * c_ij);
    const double mu_b = 1. / (xb * xb * xb * xb) * P->mu_res, mu_a = 1. / (xa * xa * xa * xa) * P->mu_res;
    if (P->avgproc == ORC_HARMONIC) return 2. / (mu_a / Ka + mu_b / Kb);
    return 1. / 2 * (Ka / mu_a + Kb / mu_b);
}
/* vertex mobility at (i+1/2, j+1/2) from the four surrounding cells: src/EllipticSolver.cpp:697-754 */
static double kv_vertex(const orc_params *P, double mq, int which, int i, int j,
                        double c_ij, double c_ip1j, double c_ijp1, double c_ip1jp1) {
    const double dx = DX(P), dy = DY(P);
    const double x = P->x_lb + (i + 0.5) * dx, y = P->y_lb + (j + 0.5) * dy;
    double K_ip1j, K_ij, K_ijp1, K_ip1jp1;
    if (which == 11) { K_ip1j = orc_K11(P, x + dx, y); K_ij = orc_K11(P, x, y); K_ijp1 = orc_K11(P, x, y + dy); K_ip1jp1 = orc_K11(P, x + dx, y + dy); }
    else             { K_ip1j = orc_K22(P, x + dx, y); K_ij = orc_K22(P, x, y); K_ijp1 = orc_K22(P, x, y + dy); K_ip1jp1 = orc_K22(P, x + dx, y + dy); }
#define MU_(cc) (1. / (((1. - (cc) + mq * (cc))) * ((1. - (cc) + mq * (cc))) * ((1. - (cc) + mq * (cc))) * ((1. - (cc) + mq * (cc)))) * P->mu_res)
    const double mu_ip1j = MU_(c_ip1j), mu_ij = MU_(c_ij), mu_ijp1 = MU_(c_ijp1), mu_ip1jp1 = MU_(c_ip1jp1);
#undef MU_
    if (P->avgproc == ORC_HARMONIC) return 4. / (mu_ip1j / K_ip1j + mu_ij / K_ij + mu_ijp1 / K_ijp1 + mu_ip1jp1 / K_ip1jp1);
    return 1. / 4 * (K_ip1j / mu_ip1j + K_ij / mu_ij + K_ijp1 / mu_ijp1 + K_ip1jp1 / mu_ip1jp1);
}

/* ------------------------------------------------------------------ pressure matrix */
enum { eC = 0, eW, eSW, eS, eSE, eE, eNE, eN, eNW };                       /* MGSolver.cpp:286-294 */
static const int OFF[9][2] = { {0,0}, {-1,0}, {-1,-1}, {0,-1}, {1,-1}, {1,0}, {1,1}, {0,1}, {-1,1} }; /* :164-166 */

static int entry_of(int di, int dj) {
    for (int e = 0; e < 9; e++) if (OFF[e][0] == di && OFF[e][1] == dj) return e;
    return -1;
}

/* unfolded weighted 9-point row: src/MGSolver.cpp:710-841 */
static void stencil_cell_t(const orc_params *P, const trig_t *t, double mq, const double *c, int i, int j, double coef[9]) {
#define C_(a, b) c[ci(P, (a), (b))]
    /* parallel (grid-aligned 5-point) part: :760-773 */
    const double K11_im12j = kx_face(P, mq, C_(i - 1, j), C_(i, j), i, j);
    const double K11_ip12j = kx_face(P, mq, C_(i, j), C_(i + 1, j), i + 1, j);
    const double K22_ijm12 = ky_face(P, mq, C_(i, j - 1), C_(i, j), i, j);
    const double K22_ijp12 = ky_face(P, mq, C_(i, j), C_(i, j + 1), i, j + 1);
    const double dx = t->dx, dy = t->dy;
    const double W_p = 1. / (dx * dx) * K11_im12j, E_p = 1. / (dx * dx) * K11_ip12j;
    const double S_p = 1. / (dy * dy) * K22_ijm12, N_p = 1. / (dy * dy) * K22_ijp12;
    const double C_p = -1. / (dy * dy) * (K22_ijm12 + K22_ijp12) - 1. / (dx * dx) * (K11_im12j + K11_ip12j);

    /* rotated-frame part: :777-841 */
    double M11_ip12j = 0, M11_im12j = 0, M22_ijp12 = 0, M22_ijm12 = 0, M12_ip12j = 0, M21_ijp12 = 0, M12_im12j = 0, M21_ijm12 = 0;
    if (P->nu != 1.) {
        const double K11_pp = kv_vertex(P, mq, 11, i, j, C_(i, j), C_(i + 1, j), C_(i, j + 1), C_(i + 1, j + 1));
        const double K11_mm = kv_vertex(P, mq, 11, i - 1, j - 1, C_(i - 1, j - 1), C_(i, j - 1), C_(i - 1, j), C_(i, j));
        const double K11_mp = kv_vertex(P, mq, 11, i - 1, j, C_(i - 1, j), C_(i, j), C_(i - 1, j + 1), C_(i, j + 1));
        const double K11_pm = kv_vertex(P, mq, 11, i, j - 1, C_(i, j - 1), C_(i + 1, j - 1), C_(i, j), C_(i + 1, j));
        const double K22_pp = kv_vertex(P, mq, 22, i, j, C_(i, j), C_(i + 1, j), C_(i, j + 1), C_(i + 1, j + 1));
        const double K22_mm = kv_vertex(P, mq, 22, i - 1, j - 1, C_(i - 1, j - 1), C_(i, j - 1), C_(i - 1, j), C_(i, j));
        const double K22_mp = kv_vertex(P, mq, 22, i - 1, j, C_(i - 1, j), C_(i, j), C_(i - 1, j + 1), C_(i, j + 1));
        const double K22_pm = kv_vertex(P, mq, 22, i, j - 1, C_(i, j - 1), C_(i + 1, j - 1), C_(i, j), C_(i + 1, j));
        M11_ip12j = t->secamt2 * (K11_pp * t->cosa2 + K22_pp * t->sina2);
        M11_im12j = t->secamt2 * (K11_mm * t->cosa2 + K22_mm * t->sina2);
        M22_ijp12 = t->secamt2 * (K22_mp * t->cost2 + K11_mp * t->sint2);
        M22_ijm12 = t->secamt2 * (K22_pm * t->cost2 + K11_pm * t->sint2);
        M12_ip12j = t->secamt2 * (K22_pp * t->costsina - K11_pp * t->cosasint);
        M21_ijp12 = t->secamt2 * (K22_mp * t->costsina - K11_mp * t->cosasint);
        M12_im12j = t->secamt2 * (K22_mm * t->costsina - K11_mm * t->cosasint);
        M21_ijm12 = t->secamt2 * (K22_pm * t->costsina - K11_pm * t->cosasint);
    }
    const double dxi = t->dxi, deta = t->deta;
    const double NE_d = M11_ip12j / (dxi * dxi), SW_d = M11_im12j / (dxi * dxi);
    const double NW_d = M22_ijp12 / (deta * deta), SE_d = M22_ijm12 / (deta * deta);
    const double C_d = -1. / (dxi * dxi) * (M11_ip12j + M11_im12j) - 1. / (deta * deta) * (M22_ijp12 + M22_ijm12);
    const double N_d = 1. / (deta * dxi) * (M12_ip12j + M21_ijp12);
    const double W_d = -1. / (deta * dxi) * (M21_ijp12 + M12_im12j);
    const double E_d = -1. / (deta * dxi) * (M12_ip12j + M21_ijm12);
    const double S_d = 1. / (deta * dxi) * (M12_im12j + M21_ijm12);
    const double nu = P->nu;                                    /* blend: :748-756 */
    coef[eC] = nu * C_p + (1 - nu) * C_d;
    coef[eW] = nu * W_p + (1 - nu) * W_d;   coef[eSW] = nu * 0. + (1 - nu) * SW_d;
    coef[eS] = nu * S_p + (1 - nu) * S_d;   coef[eSE] = nu * 0. + (1 - nu) * SE_d;
    coef[eE] = nu * E_p + (1 - nu) * E_d;   coef[eNE] = nu * 0. + (1 - nu) * NE_d;
    coef[eN] = nu * N_p + (1 - nu) * N_d;   coef[eNW] = nu * 0. + (1 - nu) * NW_d;
#undef C_
}

void orc_stencil_cell(const orc_params *P, const double *c, int i, int j, double coef[9]) {
    trig_t t = make_trig(P);
    stencil_cell_t(P, &t, pow(P->mu_res / P->mu_inj, 0.25), c, i, j, coef);
}

/* Neumann folding: every entry whose neighbour lies outside the domain is added to the entry of
 * that neighbour's mirror image and zeroed.  This one rule reproduces all eight hand-written edge
 * and corner cases of src/MGSolver.cpp:305-515 (e.g. bottom edge C+=S, W+=SW, E+=SE). */
static void fold_row(int N1, int N2, int i, int j, double coef[9]) {
    double out[9];
    for (int e = 0; e < 9; e++) out[e] = 0.;
    /* Accumulate in the reference's WRITTEN order (floating-point addition is not associative): the only sums of more than two
     * terms are the four corner diagonals, C+W+S+SW (:431, 1190), C+W+N+NW (:454, 1219), C+N+NE+E (:478, 1249), C+S+E+SE
     * (:501, 1277).  Checked bit for bit against the compiled reference (tests/test_oracle_vs_ref.py). */
    static const int order_lo[9] = { eC, eW, eE, eS, eN, eSW, eSE, eNW, eNE };      /* corners (0,0) and (0,N2-1), all edges */
    static const int order_hh[9] = { eC, eN, eNE, eE, eW, eS, eSW, eSE, eNW };      /* corner (N1-1,N2-1) */
    static const int order_hl[9] = { eC, eS, eE, eSE, eW, eN, eSW, eNW, eNE };      /* corner (N1-1,0)    */
    const int *order = (i == N1 - 1 && j == N2 - 1) ? order_hh : ((i == N1 - 1 && j == 0) ? order_hl : order_lo);
    for (int k = 0; k < 9; k++) {
        const int e = order[k];
        int ii = i + OFF[e][0], jj = j + OFF[e][1];
        if (ii < 0) ii = 0; if (ii > N1 - 1) ii = N1 - 1;
        if (jj < 0) jj = 0; if (jj > N2 - 1) jj = N2 - 1;
        out[entry_of(ii - i, jj - j)] += coef[e];
    }
    for (int e = 0; e < 9; e++) coef[e] = out[e];
}

/* test right-hand sides: src/EllipticSolver.cpp:319-353 */
static double q_test(const orc_params *P, double x, double y) {
    const double t = 2.;
    switch (P->testno) {
    case 1: return 2. / P->mu_res * cos(x) * cos(y);
    case 2: {
        const double c1 = -0.3, c2 = -0.3, a1 = exp(c1 * x), a2 = exp(c2 * y), a1_x = c1 * exp(c1 * x), a2_y = c2 * exp(c2 * y);
        return 1. / P->mu_res * ((a1 + a2) * cos(x) * cos(y) + a1_x * sin(x) * cos(y) + a2_y * cos(x) * sin(y));
    }
    case 3: {
        const double a1 = 2. + sin(t * x) * sin(y), a2 = 2. + sin(t * x) * sin(y), a1_x = t * cos(t * x) * sin(y), a2_y = cos(y) * sin(t * x);
        return 1. / P->mu_res * ((a1 + a2) * cos(x) * cos(y) + a1_x * sin(x) * cos(y) + a2_y * cos(x) * sin(y));
    }
    case 4: {
        const double a1 = 2. + sin(t * x) * sin(y), a2 = 2. + sin(x) * sin(t * y), a1_x = t * cos(t * x) * sin(y), a2_y = t * cos(t * y) * sin(x);
        return 1. / P->mu_res * ((a1 + a2) * cos(x) * cos(y) + a1_x * sin(x) * cos(y) + a2_y * cos(x) * sin(y));
    }
    default: return NAN;
    }
}

/* Dirichlet (pinned) cells of the elliptic problem: src/MGSolver.cpp:109-140 */
static int elliptic_pins(const orc_params *P, int **pi, int **pj) {
    int n = (P->stype == ORC_HORIZONTAL) ? P->N2 : 1;
    *pi = (int *)malloc(sizeof(int) * n); *pj = (int *)malloc(sizeof(int) * n);
    if (P->stype == ORC_HORIZONTAL) for (int k = 0; k < n; k++) { (*pi)[k] = 1; (*pj)[k] = k; }
    else { (*pi)[0] = 1; (*pj)[0] = 1; }
    return n;
}

/* src/MGSolver.cpp:268-570 (matrix) and :573-590 (b = -q) */
void orc_assemble_elliptic(const orc_params *P, const double *c, const double *q, double *A, double *b) {
    const int N1 = P->N1, N2 = P->N2;
    const trig_t t = make_trig(P);
    const double mq = pow(P->mu_res / P->mu_inj, 0.25);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++) {
            double coef[9];
            stencil_cell_t(P, &t, mq, c, i, j, coef);
            if (i == 0 || j == 0 || i == N1 - 1 || j == N2 - 1) fold_row(N1, N2, i, j, coef);
            memcpy(A + ((size_t)i * N2 + j) * 9, coef, sizeof(coef));
            b[(size_t)i * N2 + j] = -q[(size_t)i * N2 + j];
        }
    int *pi, *pj; const int np = elliptic_pins(P, &pi, &pj);
    for (int k = 0; k < np; k++) {                                   /* :517-567 */
        const int di = pi[k], dj = pj[k];
        if (di >= N1 || dj >= N2) continue;
        double kappa = 1.;
        if (P->is_test) {
            const double x = P->x_lb + (di + 0.5) * t.dx, y = P->y_lb + (dj + 0.5) * t.dy;
            kappa = -t.dA * q_test(P, x, y) / (cos(x) * cos(y));
        }
        double *row = A + ((size_t)di * N2 + dj) * 9;
        for (int e = 0; e < 9; e++) row[e] = 0.;
        row[eC] = kappa / t.dA;
    }
    free(pi); free(pj);
}

void orc_matvec9(int N1, int N2, const double *A, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++) {
            const double *row = A + ((size_t)i * N2 + j) * 9;
            double s = 0.;
            for (int e = 0; e < 9; e++) {
                if (row[e] == 0.) continue;
                const int ii = i + OFF[e][0], jj = j + OFF[e][1];
                s += row[e] * x[(size_t)ii * N2 + jj];
            }
            y[(size_t)i * N2 + j] = s;
        }
}

/* ------------------------------------------------------------------ multigrid-PCG
 * Stands where HYPRE StructPCG + StructSMG stand in the reference (src/MGSolver.cpp:594-629).
 * SPD form:  (S x)(i,j) = diag*x(i,j) - sum_links link*x(nb),   S = -A (elliptic) or A (parabolic).
 * Links: AW(i,j) couples (i-1,j)-(i,j); AS(i,j) couples (i,j-1)-(i,j); AV(i,j) sits on vertex
 * (i-1/2,j-1/2) and couples both (i-1,j-1)-(i,j) and (i-1,j)-(i,j-1).  Boundary links are zero
 * (that is what the Neumann folding produces). */
typedef struct level_t {
    int N1, N2;
    double *AW, *AS, *AV;      /* (N1+1)*N2, N1*(N2+1), (N1+1)*(N2+1) */
    double *diag;              /* N1*N2 (includes mass and sigma)      */
    unsigned char *fixed;      /* level 0 only: Dirichlet rows         */
    double *p, *f, *t;         /* work vectors (levels >= 1), scratch  */
} level_t;

#define LW(L, i, j) (L)->AW[(size_t)(i) * (L)->N2 + (j)]
#define LS(L, i, j) (L)->AS[(size_t)(i) * ((L)->N2 + 1) + (j)]
#define LV(L, i, j) (L)->AV[(size_t)(i) * ((L)->N2 + 1) + (j)]

static void level_free(level_t *L) { free(L->AW); free(L->AS); free(L->AV); free(L->diag); free(L->fixed); free(L->p); free(L->f); free(L->t); }

static inline double off_sum(const level_t *L, const double *x, int i, int j) {
    const int N1 = L->N1, N2 = L->N2;
    double s = 0.;
    /* fixed order W,E,S,N,SW,SE,NW,NE -- the CUDA kernels use the same order */
    if (i > 0)                     s += LW(L, i, j) * x[(size_t)(i - 1) * N2 + j];
    if (i < N1 - 1)                s += LW(L, i + 1, j) * x[(size_t)(i + 1) * N2 + j];
    if (j > 0)                     s += LS(L, i, j) * x[(size_t)i * N2 + j - 1];
    if (j < N2 - 1)                s += LS(L, i, j + 1) * x[(size_t)i * N2 + j + 1];
    if (i > 0 && j > 0)            s += LV(L, i, j) * x[(size_t)(i - 1) * N2 + j - 1];
    if (i < N1 - 1 && j > 0)       s += LV(L, i + 1, j) * x[(size_t)(i + 1) * N2 + j - 1];
    if (i > 0 && j < N2 - 1)       s += LV(L, i, j + 1) * x[(size_t)(i - 1) * N2 + j + 1];
    if (i < N1 - 1 && j < N2 - 1)  s += LV(L, i + 1, j + 1) * x[(size_t)(i + 1) * N2 + j + 1];
    return s;
}

static void level_matvec(const level_t *L, const double *x, double *y) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < L->N1; i++)
        for (int j = 0; j < L->N2; j++) {
            const size_t k = (size_t)i * L->N2 + j;
            y[k] = (L->fixed && L->fixed[k]) ? 0. : L->diag[k] * x[k] - off_sum(L, x, i, j);
        }
}

/* one colour pass of the red-black Gauss-Seidel smoother.  With the 9-point operator the diagonal
 * neighbours have the same colour; they are taken from the values before this pass (Jacobi inside a
 * colour), so a pass is order independent: it reads x and writes y (other colour copied). */
static void level_color_pass(const level_t *L, int color, const double *f, const double *x, double *y, int zero_x) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < L->N1; i++)
        for (int j = 0; j < L->N2; j++) {
            const size_t k = (size_t)i * L->N2 + j;
            if (((i + j) & 1) != color || (L->fixed && L->fixed[k])) { y[k] = zero_x ? 0. : x[k]; continue; }
            y[k] = zero_x ? f[k] / L->diag[k] : (f[k] + off_sum(L, x, i, j)) / L->diag[k];
        }
}

/* forward sweep: red (i+j even) then black; backward sweep: black then red.  Result back in x. */
static void level_smooth(level_t *L, const double *f, double *x, int forward, int zero_x) {
    const int c0 = forward ? 0 : 1;
    level_color_pass(L, c0, f, x, L->t, zero_x);
    level_color_pass(L, 1 - c0, f, L->t, x, 0);
}

/* full-weighting restriction = (bilinear prolongation)^T / 4, Neumann consistent (mirror ghosts) */
static void restrict_residual(const level_t *Lf, const double *r, const level_t *Lc, double *fc) {
    const int n1 = Lf->N1, n2 = Lf->N2;
    static const double w[4] = { 0.25, 0.75, 0.75, 0.25 };
#pragma omp parallel for schedule(static)
    for (int I = 0; I < Lc->N1; I++)
        for (int J = 0; J < Lc->N2; J++) {
            double t4[4];
            for (int a = 0; a < 4; a++) {
                int i = 2 * I - 1 + a; if (i < 0) i = 0; if (i > n1 - 1) i = n1 - 1;
                double s = 0.;
                for (int b = 0; b < 4; b++) {
                    int j = 2 * J - 1 + b; if (j < 0) j = 0; if (j > n2 - 1) j = n2 - 1;
                    s += w[b] * r[(size_t)i * n2 + j];
                }
                t4[a] = s;
            }
            fc[(size_t)I * Lc->N2 + J] = 0.25 * (w[0] * t4[0] + w[1] * t4[1] + w[2] * t4[2] + w[3] * t4[3]);
        }
}

/* cell-centred bilinear prolongation with edge replication, added to x */
static void prolong_add(const level_t *Lc, const double *e, const level_t *Lf, double *x) {
    const int n1c = Lc->N1, n2c = Lc->N2;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < Lf->N1; i++)
        for (int j = 0; j < Lf->N2; j++) {
            const size_t k = (size_t)i * Lf->N2 + j;
            if (Lf->fixed && Lf->fixed[k]) continue;
            const int I = i >> 1, J = j >> 1;
            int In = (i & 1) ? I + 1 : I - 1; if (In < 0) In = 0; if (In > n1c - 1) In = n1c - 1;
            int Jn = (j & 1) ? J + 1 : J - 1; if (Jn < 0) Jn = 0; if (Jn > n2c - 1) Jn = n2c - 1;
            const double a = 0.75 * e[(size_t)I * n2c + J] + 0.25 * e[(size_t)In * n2c + J];
            const double b = 0.75 * e[(size_t)I * n2c + Jn] + 0.25 * e[(size_t)In * n2c + Jn];
            x[k] += 0.75 * a + 0.25 * b;
        }
}

typedef struct mg_t { int nl; level_t *L; int ncoarse, w_from, w_to; } mg_t;

static void mg_cycle(mg_t *M, int l, const double *f, double *x) {
    level_t *L = &M->L[l];
    const size_t n = (size_t)L->N1 * L->N2;
    if (l == M->nl - 1) {
        for (int s = 0; s < M->ncoarse; s++) { level_smooth(L, f, x, 1, s == 0); level_smooth(L, f, x, 0, 0); }
        return;
    }
    level_t *Lc = &M->L[l + 1];
    const int gamma = (l >= M->w_from && l <= M->w_to) ? 2 : 1;
    double *r = (double *)malloc(sizeof(double) * n);
    level_smooth(L, f, x, 1, 1);                                    /* pre-smooth from a zero guess */
    for (int gcyc = 0; gcyc < gamma; gcyc++) {
        level_matvec(L, x, r);
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; k++) r[k] = (L->fixed && L->fixed[k]) ? 0. : f[k] - r[k];
        restrict_residual(L, r, Lc, Lc->f);
        mg_cycle(M, l + 1, Lc->f, Lc->p);
        prolong_add(Lc, Lc->p, L, x);
        if (gcyc < gamma - 1) { level_smooth(L, f, x, 0, 0); level_smooth(L, f, x, 1, 0); }
    }
    level_smooth(L, f, x, 0, 0);                                    /* post-smooth, reversed colour order */
    free(r);
}

static int mg_build(mg_t *M, const orc_params *P, int N1, int N2, const double *A, int sign) {
    memset(M, 0, sizeof(*M));
    M->ncoarse = P->mg_ncoarse > 0 ? P->mg_ncoarse : 8; M->w_from = P->mg_w_from; M->w_to = P->mg_w_to;
    int nl = 1, a = N1, b = N2;
    while ((a % 2 == 0) && (b % 2 == 0) && (a < b ? a : b) > (P->mg_nmin > 1 ? P->mg_nmin : 2) && (P->mg_max_levels <= 0 || nl < P->mg_max_levels)) { a /= 2; b /= 2; nl++; }
    M->nl = nl; M->L = (level_t *)calloc(nl, sizeof(level_t));
    level_t *L = &M->L[0];
    L->N1 = N1; L->N2 = N2;
    L->AW = (double *)calloc((size_t)(N1 + 1) * N2, sizeof(double));
    L->AS = (double *)calloc((size_t)N1 * (N2 + 1), sizeof(double));
    L->AV = (double *)calloc((size_t)(N1 + 1) * (N2 + 1), sizeof(double));
    L->diag = (double *)malloc(sizeof(double) * N1 * N2);
    L->fixed = (unsigned char *)calloc((size_t)N1 * N2, 1);
    L->t = (double *)malloc(sizeof(double) * N1 * N2);
    const double sg = (sign < 0) ? 1. : -1.;            /* link = +offdiag of A (elliptic) or -offdiag (parabolic) */
    double scale = 0.; int bad = 0;
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++) {
            const double *row = A + ((size_t)i * N2 + j) * 9;
            int allzero = 1;
            for (int e = 1; e < 9; e++) if (row[e] != 0.) allzero = 0;
            L->fixed[(size_t)i * N2 + j] = (unsigned char)allzero;
            L->diag[(size_t)i * N2 + j] = -sg * row[eC];
            if (fabs(row[eC]) > scale) scale = fabs(row[eC]);
        }
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++) {
            const size_t k = (size_t)i * N2 + j;
            const double *row = A + k * 9;
            if (L->fixed[k]) {                               /* take the links from the neighbours' rows */
                if (i > 0) LW(L, i, j) = sg * A[(k - N2) * 9 + eE];
                if (j > 0) LS(L, i, j) = sg * A[(k - 1) * 9 + eN];
                if (i > 0 && j > 0) LV(L, i, j) = sg * A[(k - N2 - 1) * 9 + eNE];
                continue;
            }
            if (i > 0) { LW(L, i, j) = sg * row[eW]; if (!L->fixed[k - N2] && fabs(row[eW] - A[(k - N2) * 9 + eE]) > 1e-9 * scale) bad++; }
            if (j > 0) { LS(L, i, j) = sg * row[eS]; if (!L->fixed[k - 1] && fabs(row[eS] - A[(k - 1) * 9 + eN]) > 1e-9 * scale) bad++; }
            if (i > 0 && j > 0) { LV(L, i, j) = sg * row[eSW]; if (!L->fixed[k - N2 - 1] && fabs(row[eSW] - A[(k - N2 - 1) * 9 + eNE]) > 1e-9 * scale) bad++; }
            if (i < N1 - 1 && j > 0) {                       /* SE link must sit on the same vertex value */
                const size_t kk = (size_t)(i + 1) * N2 + j;  /* cell (i+1,j): its SW vertex is (i+1/2,j-1/2) */
                if (!L->fixed[kk] && fabs(row[eSE] - A[kk * 9 + eSW]) > 1e-9 * scale) bad++;
            }
        }
    /* a fixed cell's SE/NW couplings: vertex link taken from cell (i,j)'s own SW of the NE cell etc. is covered above */
    if (bad) { fprintf(stderr, "orc mg_build: matrix is not link-symmetric (%d entries)\n", bad); return -1; }
    const double mass = (sign < 0) ? 0. : 1.;
    /* coarse levels: rediscretisation on averaged links; the pin becomes a diagonal shift sigma */
    double *sigma = NULL;
    for (int l = 1; l < nl; l++) {
        level_t *F = &M->L[l - 1], *C = &M->L[l];
        const int n1 = F->N1 / 2, n2 = F->N2 / 2;
        C->N1 = n1; C->N2 = n2;
        C->AW = (double *)calloc((size_t)(n1 + 1) * n2, sizeof(double));
        C->AS = (double *)calloc((size_t)n1 * (n2 + 1), sizeof(double));
        C->AV = (double *)calloc((size_t)(n1 + 1) * (n2 + 1), sizeof(double));
        C->diag = (double *)malloc(sizeof(double) * n1 * n2);
        C->p = (double *)calloc((size_t)n1 * n2, sizeof(double));
        C->f = (double *)calloc((size_t)n1 * n2, sizeof(double));
        C->t = (double *)malloc(sizeof(double) * n1 * n2);
        for (int I = 0; I <= n1; I++) for (int J = 0; J < n2; J++) LW(C, I, J) = (I == n1) ? 0. : 0.125 * (LW(F, 2 * I, 2 * J) + LW(F, 2 * I, 2 * J + 1));
        for (int I = 0; I < n1; I++) for (int J = 0; J <= n2; J++) LS(C, I, J) = (J == n2) ? 0. : 0.125 * (LS(F, 2 * I, 2 * J) + LS(F, 2 * I + 1, 2 * J));
        for (int I = 0; I <= n1; I++) for (int J = 0; J <= n2; J++) LV(C, I, J) = (I == n1 || J == n2) ? 0. : 0.25 * LV(F, 2 * I, 2 * J);
        double *sg_c = (double *)calloc((size_t)n1 * n2, sizeof(double));
        if (l == 1) {
            for (int i = 0; i < F->N1; i++) for (int j = 0; j < F->N2; j++) {
                if (!F->fixed[(size_t)i * F->N2 + j]) continue;
                const int bi = (i >> 1) << 1, bj = (j >> 1) << 1;
                double tot = 0.;
                for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
                    if (!di && !dj) continue;
                    const int ii = i + di, jj = j + dj;
                    if (ii < bi || ii > bi + 1 || jj < bj || jj > bj + 1) continue;
                    if (F->fixed[(size_t)ii * F->N2 + jj]) continue;
                    if (dj == 0) tot += LW(F, i > ii ? i : ii, j);
                    else if (di == 0) tot += LS(F, i, j > jj ? j : jj);
                    else tot += LV(F, i > ii ? i : ii, j > jj ? j : jj);
                }
                sg_c[(size_t)(i >> 1) * n2 + (j >> 1)] += 0.125 * tot;
            }
        } else {
            for (int I = 0; I < n1; I++) for (int J = 0; J < n2; J++)
                sg_c[(size_t)I * n2 + J] = 0.25 * (sigma[(size_t)(2 * I) * F->N2 + 2 * J] + sigma[(size_t)(2 * I + 1) * F->N2 + 2 * J]
                                                   + sigma[(size_t)(2 * I) * F->N2 + 2 * J + 1] + sigma[(size_t)(2 * I + 1) * F->N2 + 2 * J + 1]);
        }
        free(sigma); sigma = sg_c;
        for (int I = 0; I < n1; I++) for (int J = 0; J < n2; J++)
            C->diag[(size_t)I * n2 + J] = mass + sigma[(size_t)I * n2 + J]
                + (((LW(C, I, J) + LW(C, I + 1, J)) + (LS(C, I, J) + LS(C, I, J + 1)))
                   + ((LV(C, I, J) + LV(C, I + 1, J)) + (LV(C, I, J + 1) + LV(C, I + 1, J + 1))));
    }
    free(sigma);
    return 0;
}

static void mg_free(mg_t *M) { for (int l = 0; l < M->nl; l++) level_free(&M->L[l]); free(M->L); }

static double dot(const double *a, const double *b, size_t n) {
    double s = 0.;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (size_t k = 0; k < n; k++) s += a[k] * b[k];
    return s;
}

int orc_mg_apply(const orc_params *P, int N1, int N2, const double *A, int sign, const double *r, double *z) {
    mg_t M;
    if (mg_build(&M, P, N1, N2, A, sign)) return -1;
    const size_t n = (size_t)N1 * N2;
    double *f = (double *)malloc(sizeof(double) * n);
    for (size_t k = 0; k < n; k++) { f[k] = M.L[0].fixed[k] ? 0. : r[k]; z[k] = 0.; }
    mg_cycle(&M, 0, f, z);
    free(f); mg_free(&M);
    return 0;
}

/* PCG as in HYPRE's struct PCG with TwoNorm=1, RelChange=0, zero initial guess (MGSolver.cpp:577,
 * 597-618): stop when <r,r>/<b,b> < tol^2 with r the recurrence residual. */
int orc_mgpcg(const orc_params *P, int N1, int N2, const double *A, const double *b, int sign,
              double tol, double *x, int *its, double *relres, double *true_relres) {
    mg_t M;
    if (mg_build(&M, P, N1, N2, A, sign)) return -1;
    level_t *L = &M.L[0];
    const size_t n = (size_t)N1 * N2;
    double *r = (double *)malloc(sizeof(double) * n), *z = (double *)calloc(n, sizeof(double));
    double *d = (double *)malloc(sizeof(double) * n), *s = (double *)malloc(sizeof(double) * n);
    const double sg = (sign < 0) ? -1. : 1.;            /* SPD right-hand side = sg*b */
    const double bi_prod = dot(b, b, n);
    int it = 0, ret = 0; double i_prod = 0.;
    for (size_t k = 0; k < n; k++) x[k] = L->fixed[k] ? b[k] / A[k * 9 + eC] : 0.;
    if (bi_prod > 0.) {
        level_matvec(L, x, s);
        /* fixed cells couple into their neighbours through the links: off_sum reads x there */
        for (size_t k = 0; k < n; k++) r[k] = L->fixed[k] ? 0. : sg * b[k] - s[k];
        for (size_t k = 0; k < n; k++) z[k] = 0.;
        mg_cycle(&M, 0, r, z);
        memcpy(d, z, sizeof(double) * n);
        double gamma = dot(r, z, n);
        i_prod = dot(r, r, n);
        ret = 1;
        if (i_prod / bi_prod < tol * tol) ret = 0;
        while (ret && it < P->max_iter) {
            it++;
            level_matvec(L, d, s);
            const double alpha = gamma / dot(s, d, n);
#pragma omp parallel for schedule(static)
            for (size_t k = 0; k < n; k++) { x[k] += alpha * d[k]; r[k] -= alpha * s[k]; }
            i_prod = dot(r, r, n);
            if (i_prod / bi_prod < tol * tol) { ret = 0; break; }
            for (size_t k = 0; k < n; k++) z[k] = 0.;
            mg_cycle(&M, 0, r, z);
            const double gnew = dot(r, z, n), beta = gnew / gamma;
            gamma = gnew;
#pragma omp parallel for schedule(static)
            for (size_t k = 0; k < n; k++) d[k] = z[k] + beta * d[k];
        }
    }
    if (its) *its = it;
    if (relres) *relres = bi_prod > 0. ? sqrt(i_prod / bi_prod) : 0.;
    if (true_relres) {                                   /* ||b - A x|| / ||b|| with the reference's own matrix */
        orc_matvec9(N1, N2, A, x, s);
        double tt = 0.;
        for (size_t k = 0; k < n; k++) tt += (b[k] - s[k]) * (b[k] - s[k]);
        *true_relres = bi_prod > 0. ? sqrt(tt / bi_prod) : 0.;
    }
    free(r); free(z); free(d); free(s); mg_free(&M);
    return ret;
}

/* a8-a11: assemble, solve, copy into the ghosted pressure array (MGSolver.cpp:223-248, 670-706) */
int orc_solve_elliptic(const orc_params *P, orc_state *S, const double *c, const double *q, double *p) {
    const int N1 = P->N1, N2 = P->N2;
    const size_t n = (size_t)N1 * N2;
    double *A = (double *)malloc(sizeof(double) * n * 9), *b = (double *)malloc(sizeof(double) * n), *x = (double *)malloc(sizeof(double) * n);
    orc_assemble_elliptic(P, c, q, A, b);
    int its = 0; double rr = 0., tr = 0.;
    const int ret = orc_mgpcg(P, N1, N2, A, b, -1, P->eps, x, &its, &rr, &tr);
    for (int i = 0; i < N1; i++) for (int j = 0; j < N2; j++) p[ci(P, i, j)] = x[(size_t)i * N2 + j];
    if (S) { S->last_its = its; S->last_relres = rr; S->last_true_relres = tr; }
    free(A); free(b); free(x);
    return ret;
}

/* ------------------------------------------------------------------ Darcy velocity (a12) */
/* rotated-frame velocity at vertex (i-1/2,j-1/2), already rotated back to (x,y):
 * src/EllipticSolver.cpp:554-592 */
static void vertex_velocity(const orc_params *P, const trig_t *t, double mq, const double *c, const double *p, int i, int j, double *u1, double *u2) {
#define C_(a, b) c[ci(P, (a), (b))]
#define P_(a, b) p[ci(P, (a), (b))]
    const double d1dp = 1. / t->dxi * (P_(i, j) - P_(i - 1, j - 1));
    const double d2dp = 1. / t->deta * (P_(i - 1, j) - P_(i, j - 1));
    const double K11 = kv_vertex(P, mq, 11, i - 1, j - 1, C_(i - 1, j - 1), C_(i, j - 1), C_(i - 1, j), C_(i, j));
    const double K22 = kv_vertex(P, mq, 22, i - 1, j - 1, C_(i - 1, j - 1), C_(i, j - 1), C_(i - 1, j), C_(i, j));
    const double M11 = t->secamt2 * (K11 * t->cosa2 + K22 * t->sina2);
    const double M12 = t->secamt2 * (K22 * t->costsina - K11 * t->cosasint);
    const double M21 = t->secamt2 * (K22 * t->costsina - K11 * t->cosasint);
    const double M22 = t->secamt2 * (K11 * t->sint2 + K22 * t->cost2);
    const double w1 = -d1dp * M11 - d2dp * M12;
    const double w2 = -d1dp * M21 - d2dp * M22;
    *u1 = t->cost * w1 - t->sina * w2;
    *u2 = t->sint * w1 + t->cosa * w2;
#undef C_
#undef P_
}

/* src/EllipticSolver.cpp:660-683 */
void orc_calc_velocities(const orc_params *P, double *p, const double *c, double *u, double *v) {
    const int N1 = P->N1, N2 = P->N2;
    const trig_t t = make_trig(P);
    const double mq = pow(P->mu_res / P->mu_inj, 0.25), nu = P->nu;
    orc_fill_neumann_scalar(p, N1, N2, P->g);
#pragma omp parallel for schedule(static)
    for (int i = 0; i <= N1; i++)
        for (int j = 0; j < N2; j++) {
            const double par = -kx_face(P, mq, c[ci(P, i - 1, j)], c[ci(P, i, j)], i, j) * (p[ci(P, i, j)] - p[ci(P, i - 1, j)]) / t.dx;
            double a1, a2, b1, b2;
            vertex_velocity(P, &t, mq, c, p, i, j, &a1, &a2);
            vertex_velocity(P, &t, mq, c, p, i, j + 1, &b1, &b2);
            u[vi(P, i, j)] = nu * par + (1. - nu) * (0.5 * (b1 + a1));
        }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N1; i++)
        for (int j = 0; j <= N2; j++) {
            const double par = -ky_face(P, mq, c[ci(P, i, j - 1)], c[ci(P, i, j)], i, j) * (p[ci(P, i, j)] - p[ci(P, i, j - 1)]) / t.dy;
            double a1, a2, b1, b2;
            vertex_velocity(P, &t, mq, c, p, i, j, &a1, &a2);
            vertex_velocity(P, &t, mq, c, p, i + 1, j, &b1, &b2);
            v[vi(P, i, j)] = nu * par + (1. - nu) * (0.5 * (a2 + b2));
        }
    orc_fill_neumann_velocity(u, v, N1, N2, P->g);
}

/* ------------------------------------------------------------------ time step (a13) */
/* src/HyperbolicSolver.cpp:100-121 */
double orc_max_stable_advection_dt(const orc_params *P, const double *u, const double *v) {
    const double dx = DX(P), dy = DY(P);
    double dt_max = DBL_MAX;
    for (int i = 0; i < P->N1; i++)
        for (int j = 0; j < P->N2; j++) {
            double uc = fabs(0.5 * (u[vi(P, i, j)] + u[vi(P, i + 1, j)]));
            double vc = fabs(0.5 * (v[vi(P, i, j)] + v[vi(P, i, j + 1)]));
            if (uc < DBL_MIN) uc = DBL_MIN;
            if (vc < DBL_MIN) vc = DBL_MIN;
            const double a = dx / uc, b = dy / vc;
            const double dt = b < a ? b : a;
            if (dt < dt_max) dt_max = dt;
        }
    return dt_max * P->cfl;
}

/* src/HyperbolicSolver.cpp:123-198 */
double orc_max_stable_dt_combined(const orc_params *P, const double *u, const double *v) {
    const double dx = DX(P), dy = DY(P);
    double max_vel = 0.;
#pragma omp parallel for reduction(max : max_vel) schedule(static)
    for (int i = 0; i < P->N1; i++)
        for (int j = 0; j < P->N2; j++) {
            double uc = fabs(0.5 * (u[vi(P, i, j)] + u[vi(P, i + 1, j)]));
            double vc = fabs(0.5 * (v[vi(P, i, j)] + v[vi(P, i, j + 1)]));
            if (uc < DBL_MIN) uc = DBL_MIN;
            if (vc < DBL_MIN) vc = DBL_MIN;
            if (max_vel < uc) max_vel = uc;
            if (max_vel < vc) max_vel = vc;
        }
    const double a = dx / max_vel, b = dy / max_vel;
    double dt_adv = (b < a ? b : a) * P->cfl;
    double dt_diff = DBL_MAX;
    if (P->diffusion) {
        double mn = 1.e10;
#pragma omp parallel for reduction(min : mn) schedule(static)
        for (int i = 0; i < P->N1; i++)
            for (int j = 0; j < P->N2; j++) {
                const double vij = 0.5 * (v[vi(P, i, j)] + v[vi(P, i, j + 1)]);
                const double uij = 0.5 * (u[vi(P, i, j)] + u[vi(P, i + 1, j)]);
                const double v2 = vij * vij, u2 = uij * uij, absu = sqrt(v2 + u2);
                const double d11 = dx / absu * (P->alpha_l * u2 + P->alpha_t * v2);
                const double d22 = dx / absu * (P->alpha_l * v2 + P->alpha_t * u2);
                const double dt = 1. / 2 * dx * dx / (d11 + d22);
                if (dt < mn) mn = dt;
            }
        dt_diff = mn * P->dt_diff_limit;
    }
    if (P->ptype == ORC_CRANKNICOLSON) return dt_adv;
    return dt_diff < dt_adv ? dt_diff : dt_adv;
}

/* src/HyperbolicSolver.cpp:93-98 */
double orc_max_stable_dt(const orc_params *P, const double *u, const double *v) {
    return P->is_test ? orc_max_stable_advection_dt(P, u, v) : orc_max_stable_dt_combined(P, u, v);
}

/* ------------------------------------------------------------------ source step (a15) */
/* src/TRIntegrator.cpp:49-71.  Row-major serial accumulation order. */
double orc_tr_integrate(const orc_params *P, double *c, const double *q, double dt) {
    double tmo = 0.;
    for (int i = 0; i < P->N1; i++)
        for (int j = 0; j < P->N2; j++) {
            const double qij = q[(size_t)i * P->N2 + j];
            if (qij < 0) {
                double massdiff = c[ci(P, i, j)];
                c[ci(P, i, j)] = c[ci(P, i, j)] * (1 + 0.5 * dt * qij) / (1 - 0.5 * dt * qij);
                massdiff = c[ci(P, i, j)] - massdiff;
                tmo += massdiff;
            } else if (qij > 0)
                c[ci(P, i, j)] += dt * P->inj_fluid_conc * qij;
        }
    return tmo;
}

/* ------------------------------------------------------------------ fluxes (a17, a18) */
/* src/HyperbolicSolver.cpp:208-210 */
static inline double flux(const orc_params *P, double vel, double conc) { return vel * conc / P->porosity; }
/* src/UpwindSplitSolver.cpp:63-65 */
static inline double upwind_flux(const orc_params *P, double speed, double UL, double UR) { return (speed > 0) ? flux(P, speed, UL) : flux(P, speed, UR); }

/* src/MUSCLSplitSolver.cpp:61-85: the limiter argument is ignored upstream; always the van-Leer
 * form with w = 0 (algebraically minmod). */
static inline double muscl_slope(double du_m, double du_p) {
    const double w = 0.;
    double lim = 0.;
    if (du_m * du_p > 0) {
        double r = du_m / du_p;
        if (r > DBL_MAX) r = DBL_MAX;
        const double a = 2. * r / (1. + r);
        const double b = 2. * 1 / (1 - w + (1 + w) * r);
        lim = (b < a) ? b : a;                        /* std::min(a, b) */
    }
    return lim * 0.5 * ((1 + w) * du_m + (1 - w) * du_p);
}
/* src/MUSCLSplitSolver.cpp:87-93 and the half-step evolution of :122-127 */
static inline void muscl_evolve(const orc_params *P, double vL, double vR, double Um, double U0, double Up, double half_dt_over_h, double *ULbar, double *URbar) {
    const double d = muscl_slope(U0 - Um, Up - U0);
    const double UR = U0 + 0.5 * d, UL = U0 - 0.5 * d;
    const double df = half_dt_over_h * (flux(P, vL, UL) - flux(P, vR, UR));
    *URbar = UR + df; *ULbar = UL + df;
}

/* F(i,j), i in [0,N1], j in [-g,N2+g): flux through the left face of cell (i,j).
 * F is stored (N1+1) x (N2+2g), F[i*(N2+2g) + j+g].
 * MUSCL: src/MUSCLSplitSolver.cpp:102-141; upwind: src/UpwindSplitSolver.cpp:43-51 */
void orc_flux_F(const orc_params *P, int method, const double *c, const double *u, double dt, double *F) {
    const int N1 = P->N1, N2 = P->N2, g = P->g, fp = N2 + 2 * g;
    const double h = 0.5 * dt / DX(P);
#pragma omp parallel for schedule(static)
    for (int i = 0; i <= N1; i++)
        for (int j = -g; j < N2 + g; j++) {
            const double speed = u[vi(P, i, j)];
            double UL, UR;                              /* states left / right of face i-1/2 */
            if (method == ORC_UPWIND) { UL = c[ci(P, i - 1, j)]; UR = c[ci(P, i, j)]; }
            else {
                double tL, tR;
                /* left state = evolved right-extrapolated value of cell i-1 */
                muscl_evolve(P, u[vi(P, i - 1, j)], u[vi(P, i, j)], c[ci(P, i - 2, j)], c[ci(P, i - 1, j)], c[ci(P, i, j)], h, &tL, &tR);
                UL = tR;
                /* right state = evolved left-extrapolated value of cell i */
                muscl_evolve(P, u[vi(P, i, j)], u[vi(P, i + 1, j)], c[ci(P, i - 1, j)], c[ci(P, i, j)], c[ci(P, i + 1, j)], h, &tL, &tR);
                UR = tL;
            }
            F[(size_t)i * fp + (j + g)] = upwind_flux(P, speed, UL, UR);
        }
}

/* G(i,j), i in [-g,N1+g), j in [0,N2]: flux through the bottom face of cell (i,j).
 * G is stored (N1+2g) x (N2+1), G[(i+g)*(N2+1) + j].
 * MUSCL: src/MUSCLSplitSolver.cpp:152-189; upwind: src/UpwindSplitSolver.cpp:53-61 */
void orc_flux_G(const orc_params *P, int method, const double *c, const double *v, double dt, double *G) {
    const int N1 = P->N1, N2 = P->N2, g = P->g, gp = N2 + 1;
    const double h = 0.5 * dt / DY(P);
#pragma omp parallel for schedule(static)
    for (int i = -g; i < N1 + g; i++)
        for (int j = 0; j <= N2; j++) {
            const double speed = v[vi(P, i, j)];
            double UL, UR;
            if (method == ORC_UPWIND) { UL = c[ci(P, i, j - 1)]; UR = c[ci(P, i, j)]; }
            else {
                double tL, tR;
                muscl_evolve(P, v[vi(P, i, j - 1)], v[vi(P, i, j)], c[ci(P, i, j - 2)], c[ci(P, i, j - 1)], c[ci(P, i, j)], h, &tL, &tR);
                UL = tR;
                muscl_evolve(P, v[vi(P, i, j)], v[vi(P, i, j + 1)], c[ci(P, i, j - 1)], c[ci(P, i, j)], c[ci(P, i, j + 1)], h, &tL, &tR);
                UR = tL;
            }
            G[(size_t)(i + g) * gp + j] = upwind_flux(P, speed, UL, UR);
        }
}

/* src/HyperbolicSplitSolver.cpp:113-133 */
void orc_advect_x(const orc_params *P, double *c, const double *u, double dt) {
    const int N1 = P->N1, N2 = P->N2, g = P->g, fp = N2 + 2 * g;
    double *F = (double *)malloc(sizeof(double) * (size_t)(N1 + 1) * fp);
    orc_flux_F(P, P->hyper_method, c, u, dt, F);
    const double r = dt / DX(P);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N1; i++)
        for (int j = -g; j < N2 + g; j++)
            c[ci(P, i, j)] += r * (F[(size_t)i * fp + (j + g)] - F[(size_t)(i + 1) * fp + (j + g)]);
    free(F);
}
/* src/HyperbolicSplitSolver.cpp:139-159 */
void orc_advect_y(const orc_params *P, double *c, const double *v, double dt) {
    const int N1 = P->N1, N2 = P->N2, g = P->g, gp = N2 + 1;
    double *G = (double *)malloc(sizeof(double) * (size_t)(N1 + 2 * g) * gp);
    orc_flux_G(P, P->hyper_method, c, v, dt, G);
    const double r = dt / DY(P);
#pragma omp parallel for schedule(static)
    for (int i = -g; i < N1 + g; i++)
        for (int j = 0; j < N2; j++)
            c[ci(P, i, j)] += r * (G[(size_t)(i + g) * gp + j] - G[(size_t)(i + g) * gp + j + 1]);
    free(G);
}

/* src/HyperbolicSplitSolver.cpp:51-106 */
void orc_advect_split(const orc_params *P, orc_state *S, double *c, const double *u, const double *v, double dt) {
    orc_fill_boundaries(P, c);
    switch (P->dim_splitting) {
    case ORC_ALTERNATING:
        if (S->prev_first_sweep == 1) { orc_advect_x(P, c, u, dt); orc_advect_y(P, c, v, dt); S->prev_first_sweep = 0; }
        else { orc_advect_y(P, c, v, dt); orc_advect_x(P, c, u, dt); S->prev_first_sweep = 1; }
        break;
    case ORC_X_FIRST: orc_advect_x(P, c, u, dt); orc_advect_y(P, c, v, dt); S->prev_first_sweep = 0; break;
    case ORC_Y_FIRST: orc_advect_y(P, c, v, dt); orc_advect_x(P, c, u, dt); S->prev_first_sweep = 1; break;
    case ORC_STRANG:  orc_advect_x(P, c, u, 0.5 * dt); orc_advect_y(P, c, v, dt); orc_advect_x(P, c, u, 0.5 * dt); S->prev_first_sweep = 0; break;
    default: {
        const size_t n = (size_t)(P->N1 + 2 * P->g) * SP(P);
        double *ct = (double *)malloc(sizeof(double) * n);
        memcpy(ct, c, sizeof(double) * n);
        orc_advect_x(P, c, u, dt);  orc_advect_y(P, c, v, dt);
        orc_advect_y(P, ct, v, dt); orc_advect_x(P, ct, u, dt);
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; k++) c[k] = 0.5 * (c[k] + ct[k]);
        free(ct);
        S->prev_first_sweep = 0;
    } break;
    }
}

/* ------------------------------------------------------------------ dispersion (a19) */
/* src/MGSolver.cpp:846-881 */
void orc_dispersion_F(const orc_params *P, const double *u, const double *v, int i, int j, double *d11, double *d12) {
    const double dx = DX(P), dy = DY(P);
    const double uu = u[vi(P, i, j)];
    const double vv = 1. / 4 * (v[vi(P, i - 1, j + 1)] + v[vi(P, i, j + 1)] + v[vi(P, i - 1, j)] + v[vi(P, i, j)]);
    const double u2 = uu * uu, v2 = vv * vv;
    const double maxdx = dx > dy ? dx : dy;
    const double al = P->alpha_l * maxdx, at = P->alpha_t * maxdx;
    const double fl = P->porosity * al / (sqrt(u2 + v2)), ft = P->porosity * at / (sqrt(u2 + v2));
    *d11 = fl * u2 + ft * v2;
    *d12 = fl * (uu * vv) + ft * (-uu * vv);
}
/* src/MGSolver.cpp:886-919 */
void orc_dispersion_G(const orc_params *P, const double *u, const double *v, int i, int j, double *d22, double *d21) {
    const double dx = DX(P), dy = DY(P);
    const double uu = 1. / 4 * (u[vi(P, i + 1, j - 1)] + u[vi(P, i + 1, j)] + u[vi(P, i, j - 1)] + u[vi(P, i, j)]);
    const double vv = v[vi(P, i, j)];
    const double u2 = uu * uu, v2 = vv * vv;
    const double maxdx = dx > dy ? dx : dy;
    const double al = P->alpha_l * maxdx, at = P->alpha_t * maxdx;
    const double fl = P->porosity * al / (sqrt(u2 + v2)), ft = P->porosity * at / (sqrt(u2 + v2));
    *d21 = fl * (uu * vv) + ft * (-uu * vv);
    *d22 = fl * v2 + ft * u2;
}

/* Crank-Nicolson matrix and right-hand side: src/MGSolver.cpp:923-948, 997-1321, 1324-1371.
 * dt is the (half) step handed to Solve_Parabolic_cc_to_c.  No pinned cell for the parallel /
 * diagonal sources (N_para_dir = 0, MGSolver.cpp:117). */
void orc_assemble_parabolic(const orc_params *P, const double *c, const double *u, const double *v, double dt, double *A, double *b) {
    const int N1 = P->N1, N2 = P->N2;
    const double dx = DX(P), dy = DY(P);
#define C_(a, b) c[ci(P, (a), (b))]
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N1; i++)
        for (int j = 0; j < N2; j++) {
            double d11_im, d12_im, d22_jm, d21_jm, d22_jp, d21_jp, d11_ip, d12_ip;
            orc_dispersion_F(P, u, v, i, j, &d11_im, &d12_im);
            orc_dispersion_G(P, u, v, i, j, &d22_jm, &d21_jm);
            orc_dispersion_G(P, u, v, i, j + 1, &d22_jp, &d21_jp);
            orc_dispersion_F(P, u, v, i + 1, j, &d11_ip, &d12_ip);
            double k[9];
            k[eNW] = dt / (8. * dx * dy) * (d12_im + d21_jp);
            k[eNE] = -dt / (8. * dx * dy) * (d12_ip + d21_jp);
            k[eN] = -dt / 8. * (4. / (dy * dy) * d22_jp + 1. / (dx * dy) * (-d12_im + d12_ip));
            k[eSW] = -dt / (8. * dx * dy) * (d12_im + d21_jm);
            k[eSE] = dt / (8. * dx * dy) * (d12_ip + d21_jm);
            k[eS] = -dt / 8. * (4. / (dy * dy) * d22_jm + 1. / (dx * dy) * (-d12_ip + d12_im));
            k[eW] = -dt / 8. * (4. / (dx * dx) * d11_im + 1. / (dx * dy) * (-d21_jp + d21_jm));
            k[eC] = 1. + dt / (2. * dx * dx) * (d11_ip + d11_im) + dt / (2. * dy * dy) * (d22_jp + d22_jm);
            k[eE] = -dt / 8. * (4. / (dx * dx) * d11_ip + 1. / (dx * dy) * (-d21_jm + d21_jp));
            if (i == 0 || j == 0 || i == N1 - 1 || j == N2 - 1) fold_row(N1, N2, i, j, k);
            memcpy(A + ((size_t)i * N2 + j) * 9, k, sizeof(k));
            b[(size_t)i * N2 + j] = C_(i, j) + dt / 8. * (
                4. / (dx * dx) * (d11_ip * (C_(i + 1, j) - C_(i, j)) - d11_im * (C_(i, j) - C_(i - 1, j)))
                + 1. / (dx * dy) * (d12_ip * (C_(i, j + 1) + C_(i + 1, j + 1) - C_(i, j - 1) - C_(i + 1, j - 1))
                                    - d12_im * (C_(i, j + 1) + C_(i - 1, j + 1) - C_(i - 1, j - 1) - C_(i, j - 1)))
                + 1. / (dx * dy) * (d21_jp * (C_(i + 1, j) + C_(i + 1, j + 1) - C_(i - 1, j + 1) - C_(i - 1, j))
                                    - d21_jm * (C_(i + 1, j) + C_(i + 1, j - 1) - C_(i - 1, j) - C_(i - 1, j - 1)))
                + 4. / (dy * dy) * (d22_jp * (C_(i, j + 1) - C_(i, j)) - d22_jm * (C_(i, j) - C_(i, j - 1))));
        }
#undef C_
    if (P->stype == ORC_HORIZONTAL) {                               /* :1292-1315, 1357-1364 */
        const double dA = dx * dy;
        for (int j = 0; j < N2; j++) {
            const int i = N1 - 2;
            double *row = A + ((size_t)i * N2 + j) * 9;
            for (int e = 0; e < 9; e++) row[e] = 0.;
            row[eC] = 1. / dA; b[(size_t)i * N2 + j] = 0.;
        }
    }
}

/* BiCGStab with diagonal (Jacobi) preconditioning on the reference's 9-entry matrix: for the NON-symmetric Crank-Nicolson matrix of
 * alpha_l != alpha_t (src/MGSolver.cpp:939-947), to which upstream applies HYPRE's PCG regardless.  Same stop rule as everywhere:
 * ||r||_2 / ||b||_2 < tol with the recurrence residual, x0 = 0. */
int orc_bicgstab9(int N1, int N2, const double *A, const double *b, double tol, int max_iter, double *x, int *its, double *relres) {
    const size_t n = (size_t)N1 * N2;
    double *r = (double *)malloc(sizeof(double) * n), *r0 = (double *)malloc(sizeof(double) * n), *p = (double *)calloc(n, sizeof(double));
    double *v = (double *)calloc(n, sizeof(double)), *s = (double *)malloc(sizeof(double) * n), *t = (double *)malloc(sizeof(double) * n);
    double *ph = (double *)malloc(sizeof(double) * n), *sh = (double *)malloc(sizeof(double) * n);
    const double bb = dot(b, b, n);
    for (size_t k = 0; k < n; k++) { x[k] = 0.; r[k] = b[k]; r0[k] = b[k]; }
    double rho = 1., alpha = 1., omega = 1., rr = bb;
    int it = 0, ret = (bb > 0. && rr / bb >= tol * tol) ? 1 : 0;
    while (ret && it < max_iter) {
        it++;
        const double rho1 = dot(r0, r, n);
        if (rho1 == 0.) { ret = -1; break; }
        const double beta = (rho1 / rho) * (alpha / omega);
        rho = rho1;
        for (size_t k = 0; k < n; k++) { p[k] = r[k] + beta * (p[k] - omega * v[k]); ph[k] = p[k] / A[k * 9 + eC]; }
        orc_matvec9(N1, N2, A, ph, v);
        alpha = rho / dot(r0, v, n);
        for (size_t k = 0; k < n; k++) { s[k] = r[k] - alpha * v[k]; sh[k] = s[k] / A[k * 9 + eC]; }
        orc_matvec9(N1, N2, A, sh, t);
        const double tt = dot(t, t, n);
        omega = tt > 0. ? dot(t, s, n) / tt : 0.;
        for (size_t k = 0; k < n; k++) { x[k] += alpha * ph[k] + omega * sh[k]; r[k] = s[k] - omega * t[k]; }
        rr = dot(r, r, n);
        if (rr / bb < tol * tol) { ret = 0; break; }
        if (omega == 0.) { ret = -1; break; }
    }
    if (its) *its = it;
    if (relres) *relres = bb > 0. ? sqrt(rr / bb) : 0.;
    free(r); free(r0); free(p); free(v); free(s); free(t); free(ph); free(sh);
    return ret;
}

/* src/MGSolver.cpp:952-976 */
int orc_solve_parabolic(const orc_params *P, orc_state *S, double *c, const double *u, const double *v, double dt) {
    const int N1 = P->N1, N2 = P->N2;
    const size_t n = (size_t)N1 * N2;
    double *A = (double *)malloc(sizeof(double) * n * 9), *b = (double *)malloc(sizeof(double) * n), *x = (double *)malloc(sizeof(double) * n);
    orc_assemble_parabolic(P, c, u, v, dt, A, b);
    int its = 0; double rr = 0.;
    orc_params Q = *P; Q.mg_max_levels = 1; Q.mg_ncoarse = 1;       /* well conditioned: symmetric-GS PCG */
    const int ret = (P->alpha_l != P->alpha_t) ? orc_bicgstab9(N1, N2, A, b, P->para_eps, P->max_iter, x, &its, &rr)   /* non-symmetric matrix */
                                               : orc_mgpcg(&Q, N1, N2, A, b, +1, P->para_eps, x, &its, &rr, NULL);
    for (int i = 0; i < N1; i++) for (int j = 0; j < N2; j++) c[ci(P, i, j)] = x[(size_t)i * N2 + j];
    if (S) S->last_parabolic_its = its;
    free(A); free(b); free(x);
    return ret < 0 ? ret : 0;
}

/* ------------------------------------------------------------------ transport step (a14) */
/* src/HyperbolicSplitSolver.cpp:162-198 with SPLIT2ND (GenSimDefs.h:18) */
int orc_transport_step(const orc_params *P, orc_state *S, double *c, const double *u, const double *v, const double *q, double dt) {
    int ret = 0;
    S->total_mass_outflux += fabs(orc_tr_integrate(P, c, q, 0.5 * dt));     /* HyperbolicSolver.cpp:201-205 */
    orc_fill_boundaries(P, c);
    if (P->diffusion) ret |= orc_solve_parabolic(P, S, c, u, v, 0.5 * dt);
    orc_advect_split(P, S, c, u, v, dt);
    S->total_mass_outflux += fabs(orc_tr_integrate(P, c, q, 0.5 * dt));
    orc_fill_boundaries(P, c);
    if (P->diffusion) ret |= orc_solve_parabolic(P, S, c, u, v, 0.5 * dt);
    return ret;
}

/* src/Impec2DSolver.cpp:339-389 */
double orc_impes_step(const orc_params *P, orc_state *S, double *p, double *c, double *u, double *v, const double *q, double t, double t_next_output) {
    orc_fill_boundaries(P, c);
    orc_solve_elliptic(P, S, c, q, p);
    orc_calc_velocities(P, p, c, u, v);
    double dt = orc_max_stable_dt(P, u, v);
    if (t_next_output > t && t + dt >= t_next_output) dt = t_next_output - t;
    orc_transport_step(P, S, c, u, v, q, dt);
    return dt;
}

/* ------------------------------------------------------------------ wells */
/* src/Util.cpp:444-491 */
double orc_exact_vol_frac_circle(double x1, double y1, double r, double dx, double dy) {
    if (x1 * x1 + y1 * y1 < r * r) {
        const double y2 = y1 + dy, x2 = x1 + dx;
        if (x2 * x2 + y2 * y2 > r * r) {
            const double xs = sqrt(r * r - y1 * y1), ys = sqrt(r * r - x1 * x1);
            double area = 0, xub, xlb, yub;
            if (ys <= y2 && xs <= x2) { xlb = x1; xub = xs; yub = ys; }
            else if (ys < y2 && xs > x2) { yub = ys; xlb = x1; xub = x2; }
            else if (ys > y2 && xs < x2) { yub = y2; xlb = sqrt(r * r - yub * yub); xub = xs; area += (xlb - x1) * dy; }
            else { yub = y2; xlb = sqrt(r * r - yub * yub); xub = x2; area += dy * (xlb - x1); }
            (void)yub;
            const double t1 = asin(xlb / r), t2 = asin(xub / r);
            area += r * r / 2 * (sin(2 * t2) / 2. + t2 - sin(2 * t1) / 2. - t1);
            area -= y1 * (xub - xlb);
            return area / (dx * dy);
        }
        return 1.;
    }
    return 0;
}

/* src/Impec2DSolver.cpp:241-280 with src/Util.cpp:497-586.  The reference builds the quarter circle
 * in a ghosted array and reverses the WHOLE array (ghosts included); with symmetric ghost widths the
 * interior maps (i,j) -> (N1-1-i, N2-1-j). */
void orc_init_source(const orc_params *P, double I, double well_rc, int well_profile, double *q) {
    const int N1 = P->N1, N2 = P->N2;
    const double dx = DX(P), dy = DY(P);
    const double r = well_rc * (P->x_ub - P->x_lb);
    const double circval = I * 4. / (M_PI * r * r), qval = I / (dx * dy);
    const size_t M = (size_t)N1 * N2;
    for (size_t k = 0; k < M; k++) q[k] = 0.;
    if (well_profile == 2) {                                         /* one_cell: Util.cpp:565-586 */
        q[0] = qval;
        if (P->stype == ORC_PARALLEL) { q[M - 1] = qval; q[N2 - 1] = -qval; q[(size_t)(N1 - 1) * N2] = -qval; }
        else q[M - 1] = -qval;
        return;
    }
    double *qI = (double *)malloc(sizeof(double) * M);
    for (int i = 0; i < N1; i++) for (int j = 0; j < N2; j++)
        qI[(size_t)i * N2 + j] = circval * orc_exact_vol_frac_circle(P->x_lb + i * dx, P->y_lb + j * dy, r, dx, dy);
#define QI(i, j) qI[(size_t)(i) * N2 + (j)]
    for (int i = 0; i < N1; i++) for (int j = 0; j < N2; j++) {
        if (P->stype == ORC_PARALLEL)  /* q = (qI + rev01(qI)) - rev1(qI) - rev0(qI) in that order: Util.cpp:540-553 */
            q[(size_t)i * N2 + j] = ((QI(i, j) + QI(N1 - 1 - i, N2 - 1 - j)) - QI(i, N2 - 1 - j)) - QI(N1 - 1 - i, j);
        else
            q[(size_t)i * N2 + j] = QI(i, j) - QI(N1 - 1 - i, N2 - 1 - j);
    }
#undef QI
    free(qI);
}

/* ------------------------------------------------------------------ profiles and norms */
/* src/Util.cpp:55-114 */
/* InitHorizontalFront, src/Util.cpp:607-639.  mode: 0 pulsing, 1 linear, 2 constant (enum INJ_MODE, src/SimParams.h:27).  Returns the
 * layer width in rows (< 1: the reference exits). */
int orc_init_horizontal_front(const orc_params *P, double I, double well_width, int mode, double q_freq, double q_high_multiple,
                              double tstart, double tfinal, double t, double *q) {
    const int N1 = P->N1, N2 = P->N2;
    const double dx = DX(P), dy = DY(P);
    const int qlayer_width = (int)round(well_width * (P->x_ub - P->x_lb) / dx);
    if (qlayer_width < 1) return qlayer_width;
    double val = I / (N2 * dx * dy * qlayer_width);
    const size_t M = (size_t)N1 * N2;
    for (size_t i = 0; i < M; i++) q[i] = 0.;
    if (mode == 0) {
        const double tau = 1. / q_freq;
        const double x = fmod(t, tau);
        if (x > tau / 2.) val *= q_high_multiple;
    } else if (mode == 1) {
        const double vl = val, vh = val * q_high_multiple;
        val = vl + (t - tstart) * (vh - vl) / (tfinal - tstart);
    }
    for (int i = 0; i < N2 * qlayer_width; i++) { q[i] = val; q[M - 1 - i] = -val; }
    return qlayer_width;
}

double orc_profile(const orc_params *P, int profile, double x, double y) {
    const double dx = DX(P), dy = DY(P);
    switch (profile) {
    case ORC_PROFILE_SINE:
        return 1. + 1. / (dx * dy * 4. * M_PI * M_PI) * (cos(2. * M_PI * (x + dx / 2.)) - cos(2. * M_PI * (x - dx / 2.)))
                                                       * (cos(2. * M_PI * (y + dy / 2.)) - cos(2. * M_PI * (y - dy / 2.)));
    case ORC_PROFILE_TOPHAT: {
        const double width = 0.3 * (P->x_ub - P->x_lb), height = 0.3 * (P->y_ub - P->y_lb);
        const double xc = (P->x_ub - P->x_lb) / 2., yc = (P->y_ub - P->y_lb) / 2.;
        return (fabs(x - xc) <= width / 2. && fabs(y - yc) <= height / 2.) ? 1. : 0.;
    }
    case ORC_PROFILE_GAUSS: {
        const double xc = (P->x_ub - P->x_lb) / 2., yc = (P->y_ub - P->y_lb) / 2.;
        const double intx = 1. / 16 * sqrt(M_PI) * (-erf(8. * xc - 4. * dx - 8. * x) + erf(8. * xc + 4. * dx - 8. * x));
        const double inty = 1. / 16 * sqrt(M_PI) * (-erf(8. * yc - 4. * dy - 8. * y) + erf(8. * yc + 4. * dy - 8. * y));
        return 1. / (dx * dy) * intx * inty;
    }
    case ORC_PROFILE_HSINE:   /* the driver maps HSINE to VerticalSine: Impec2DSolver.cpp:867-869 */
        return 1. + 1. / (dy) * (-1. / (2. * M_PI)) * (cos(2 * M_PI * (y + dy / 2.)) - cos(2 * M_PI * (y - dy / 2.)));
    default: return NAN;
    }
}
/* src/Util.cpp:119-147 */
double orc_advection_norm(const orc_params *P, const double *c, int profile, int norm) {
    const double dx = DX(P), dy = DY(P);
    double s = 0.;
    for (int i = 0; i < P->N1; i++)
        for (int j = 0; j < P->N2; j++) {
            const double x = P->x_lb + (i + 0.5) * dx, y = P->y_lb + (j + 0.5) * dy;
            const double ex = orc_profile(P, profile, x, y);
            if (norm == 2) s += (ex - c[ci(P, i, j)]) * (ex - c[ci(P, i, j)]);
            else s += fabs(ex - c[ci(P, i, j)]);
        }
    return norm == 2 ? sqrt(dx * dy * s) : s * (dx * dy);
}

/* src/EllipticSolver.cpp:357-410 */
void orc_init_elliptic_test(const orc_params *P, double *p_ex, double *q, double *u_ex, double *v_ex) {
    const double dx = DX(P), dy = DY(P);
    for (int i = 0; i <= P->N1; i++) for (int j = 0; j < P->N2; j++) {
        const double x = P->x_lb + i * dx, y = P->y_lb + (j + 0.5) * dy;
        u_ex[vi(P, i, j)] = orc_K11(P, x, y) / P->mu_res * sin(x) * cos(y);
    }
    for (int i = 0; i < P->N1; i++) for (int j = 0; j <= P->N2; j++) {
        const double x = P->x_lb + (i + 0.5) * dx, y = P->y_lb + j * dy;
        v_ex[vi(P, i, j)] = orc_K22(P, x, y) / P->mu_res * cos(x) * sin(y);
    }
    for (int i = 0; i < P->N1; i++) for (int j = 0; j < P->N2; j++) {
        const double x = P->x_lb + (i + 0.5) * dx, y = P->y_lb + (j + 0.5) * dy;
        p_ex[ci(P, i, j)] = cos(x) * cos(y);
        q[(size_t)i * P->N2 + j] = q_test(P, x, y);
    }
}

/* ------------------------------------------------------------------ diagnostics */
/* src/Impec2DSolver.cpp:607-620 */
double orc_total_mass(const orc_params *P, const double *c) {
    double s = 0.;
    for (int i = 0; i < P->N1; i++) for (int j = 0; j < P->N2; j++) s += c[ci(P, i, j)];
    return s;
}
/* src/Impec2DSolver.cpp:542-557 (note: dx*dx, as upstream) */
double orc_divergence_error(const orc_params *P, const double *u, const double *v, const double *q) {
    const double dx = DX(P), dy = DY(P);
    double err = 0.;
    for (int i = 0; i < P->N1; i++) for (int j = 0; j < P->N2; j++) {
        const double div = (v[vi(P, i, j + 1)] - v[vi(P, i, j)]) / dy + (u[vi(P, i + 1, j)] - u[vi(P, i, j)]) / dx;
        err += fabs(div - q[(size_t)i * P->N2 + j]);
    }
    return err * dx * dx;
}
/* src/Impec2DSolver.cpp:562-588 */
double orc_boundary_velocity_error(const orc_params *P, const double *u, const double *v) {
    double s = 0.;
    for (int i = 0; i < P->N1; i++) { s += fabs(v[vi(P, i, 0)]); s += fabs(v[vi(P, i, P->N2)]); }
    for (int j = 0; j < P->N2; j++) { s += fabs(u[vi(P, 0, j)]); s += fabs(u[vi(P, P->N1, j)]); }
    return s;
}

/* ------------------------------------------------------------------ synthetic developed state */
static inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
/* ---------------------------------------------------------------- parallel -> diagonal projection
 * Impec2DSolver::Project (src/Impec2DSolver.cpp:690-716) + ParallelProjector (src/ParallelProjector.cpp:52-199): the diagonal
 * grid has round(N/sqrt 2) cells of width L/(sqrt 2 dN); its cell centre (x,y) is rotated by -pi/4 into the parallel frame,
 * y is mirrored (|y|), and the ghosted parallel field is interpolated bilinearly between the cells floor(x/L N), +1
 * (the reference does not re-centre, so the weights extrapolate in the lower half of a cell; reproduced as is). */
void orc_project_dims(const orc_params *P, int *dN1, int *dN2) {
    *dN1 = (int)round(P->N1 / sqrt(2)); *dN2 = (int)round(P->N2 / sqrt(2));
}
void orc_project_parallel(const orc_params *P, const double *a, int dN1, int dN2, double *out) {
    const double angle = -M_PI / 4.;
    const double ca = cos(angle), sa = sin(angle);
    const double p_dx = (P->x_ub - P->x_lb) / P->N1, p_dy = (P->y_ub - P->y_lb) / P->N2;
    const double d_dx = (P->x_ub - P->x_lb) / (sqrt(2.) * dN1), d_dy = (P->y_ub - P->y_lb) / (sqrt(2.) * dN2);
    for (int i = 0; i < dN1; i++) for (int j = 0; j < dN2; j++) {
        const double xd = P->x_lb + (i + 0.5) * d_dx, yd = P->y_lb + (j + 0.5) * d_dy;      /* centred() */
        const double x = ca * xd - sa * yd;                                                     /* DiagonalToParallel */
        const double y = fabs(sa * xd + ca * yd);
        const int x00 = (int)floor(x / (P->x_ub - P->x_lb) * P->N1), y00 = (int)floor(y / (P->y_ub - P->y_lb) * P->N2);
        const double x_00 = P->x_lb + (x00 + 0.5) * p_dx, x_11 = P->x_lb + (x00 + 1 + 0.5) * p_dx;
        const double y_00 = P->y_lb + (y00 + 0.5) * p_dy, y_11 = P->y_lb + (y00 + 1 + 0.5) * p_dy;
        out[(size_t)i * dN2 + j] =
            1. / ((x_11 - x_00) * (y_11 - y_00)) * (x_11 - x) * (y_11 - y) * a[ci(P, x00, y00)] +
            1. / ((x_11 - x_00) * (y_11 - y_00)) * (x - x_00) * (y_11 - y) * a[ci(P, x00 + 1, y00)] +
            1. / ((x_11 - x_00) * (y_11 - y_00)) * (x_11 - x) * (y - y_00) * a[ci(P, x00, y00 + 1)] +
            1. / ((x_11 - x_00) * (y_11 - y_00)) * (x - x_00) * (y - y_00) * a[ci(P, x00 + 1, y00 + 1)];
    }
}

/* c = 1/2 [1 - tanh((r - r0(theta))/w)], r0 = L (0.45 + 0.08 sin 16 theta), w = 4 dx,
 * plus 1e-3 (U - 1/2) counter-based noise, clipped to [0,1]; the parallel source has a second injector
 * in the opposite corner.  Interior only; caller fills ghosts. */
void orc_synth_state(const orc_params *P, unsigned long long seed, double *c) {
    const double dx = DX(P), dy = DY(P), L = P->x_ub - P->x_lb, w = 4. * dx;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P->N1; i++)
        for (int j = 0; j < P->N2; j++) {
            const double x = (i + 0.5) * dx, y = (j + 0.5) * dy;
            double r = sqrt(x * x + y * y), th = atan2(y, x);
            double val = 0.5 * (1. - tanh((r - L * (0.45 + 0.08 * sin(16. * th))) / w));
            if (P->stype == ORC_PARALLEL) {
                const double x2 = L - x, y2 = (P->y_ub - P->y_lb) - y;
                r = sqrt(x2 * x2 + y2 * y2); th = atan2(y2, x2);
                const double v2 = 0.5 * (1. - tanh((r - L * (0.45 + 0.08 * sin(16. * th))) / w));
                if (v2 > val) val = v2;
            }
            const unsigned long long h = splitmix64(seed ^ ((unsigned long long)i * (unsigned long long)P->N2 + (unsigned long long)j));
            val += 1e-3 * ((double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5);
            if (val < 0.) val = 0.; if (val > 1.) val = 1.;
            c[ci(P, i, j)] = val;
        }
}
