// host_util.inl -- host-side setup the hot path's callers need once per run (not kernels):
// well/source initialisation (src/Impec2DSolver.cpp:241-280, src/Util.cpp:444-586) and the synthetic
// developed concentration field used by bench.py and the tests.
#include <thread>

// exact area fraction of the cell [x1,x1+dx] x [y1,y1+dy] inside the circle of radius r about the origin
// (src/Util.cpp:444-491)
static double exact_vol_frac_circle(double x1, double y1, double r, double dx, double dy) {
    if (!(x1 * x1 + y1 * y1 < r * r)) return 0;
    const double y2 = y1 + dy, x2 = x1 + dx;
    if (!(x2 * x2 + y2 * y2 > r * r)) return 1.;
    const double xs = sqrt(r * r - y1 * y1), ys = sqrt(r * r - x1 * x1);
    double area = 0, xub, xlb;
    if (ys <= y2 && xs <= x2) { xlb = x1; xub = xs; }
    else if (ys < y2 && xs > x2) { xlb = x1; xub = x2; }
    else if (ys > y2 && xs < x2) { xlb = sqrt(r * r - y2 * y2); xub = xs; area += (xlb - x1) * dy; }
    else { xlb = sqrt(r * r - y2 * y2); xub = x2; area += dy * (xlb - x1); }
    const double t1 = asin(xlb / r), t2 = asin(xub / r);
    area += r * r / 2 * (sin(2 * t2) / 2. + t2 - sin(2 * t1) / 2. - t1);
    area -= y1 * (xub - xlb);
    return area / (dx * dy);
}

template <class F> static void parallel_rows(int n, F f) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1; if (nt > 64) nt = 64; if ((int)nt > n) nt = (unsigned)n;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([=]() { for (int i = (int)t; i < n; i += (int)nt) f(i); });
    for (auto &x : th) x.join();
}

// InitializeSource: quarter-circle wells (diagonal: injector at (0,0), producer opposite; parallel: injectors
// at (0,0) and (N1-1,N2-1), producers in the other two corners) or one-cell wells.
extern "C" int impec_init_source_rows(const impec_params *P, double I, double well_rc, int well_profile, int row0, int nrows, double *q);
extern "C" int impec_init_source(const impec_params *P, double I, double well_rc, int well_profile, double *q) {
    return impec_init_source_rows(P, I, well_rc, well_profile, 0, P ? P->N1 : 0, q);
}
extern "C" int impec_init_source_rows(const impec_params *P, double I, double well_rc, int well_profile, int row0, int nrows, double *q) {
    if (!P || !q || row0 < 0 || row0 + nrows > P->N1) return IMPEC_ERR_ARG;
    if (P->stype == IMPEC_HORIZONTAL) return IMPEC_ERR_UNSUPPORTED;           // see impec_init_horizontal_front
    const int N1 = P->N1, N2 = P->N2;
    const double dx = (P->x_ub - P->x_lb) / N1, dy = (P->y_ub - P->y_lb) / N2;
    const double r = well_rc * (P->x_ub - P->x_lb);
    const double circval = I * 4. / (M_PI * r * r), qval = I / (dx * dy);
    if (well_profile == 2) {                                       // one_cell (src/Util.cpp:565-586)
        for (size_t k = 0; k < (size_t)nrows * N2; k++) q[k] = 0.;
        auto put = [&](int i, int j, double v) { if (i >= row0 && i < row0 + nrows) q[(size_t)(i - row0) * N2 + j] = v; };
        put(0, 0, qval);
        if (P->stype == IMPEC_PARALLEL) { put(N1 - 1, N2 - 1, qval); put(0, N2 - 1, -qval); put(N1 - 1, 0, -qval); }
        else put(N1 - 1, N2 - 1, -qval);
        return IMPEC_OK;
    }
    if (well_profile != 0) return IMPEC_ERR_UNSUPPORTED;           // "gauss" is a no-op upstream (:253-254)
    auto qI = [&](int i, int j) { return circval * exact_vol_frac_circle(P->x_lb + i * dx, P->y_lb + j * dy, r, dx, dy); };
    const bool par = P->stype == IMPEC_PARALLEL;
    parallel_rows(nrows, [&](int il) {
        const int i = row0 + il;
        for (int j = 0; j < N2; j++) {
            const double a = qI(i, j), b = qI(N1 - 1 - i, N2 - 1 - j);
            q[(size_t)il * N2 + j] = par ? ((a + b) - qI(i, N2 - 1 - j)) - qI(N1 - 1 - i, j) : a - b;
        }
    });
    return IMPEC_OK;
}

// InitHorizontalFront (src/Util.cpp:607-639): a layer of round(well_width (x_ub - x_lb) / dx) injecting rows at i = 0.. and as many
// producing rows at the far end; mode 0 pulsing (rate * q_high_multiple in the second half of every period 1/q_freq), 1 linear ramp from
// the base rate at tstart to q_high_multiple times it at tfinal, 2 constant.  Returns IMPEC_ERR_ARG when the layer is narrower than a cell.
extern "C" int impec_init_horizontal_front(const impec_params *P, double I, double well_width, int mode, double q_freq, double q_high_multiple,
                                           double tstart, double tfinal, double t, double *q) {
    if (!P || !q) return IMPEC_ERR_ARG;
    const int N1 = P->N1, N2 = P->N2;
    const double dx = (P->x_ub - P->x_lb) / N1, dy = (P->y_ub - P->y_lb) / N2;
    const int width = (int)round(well_width * (P->x_ub - P->x_lb) / dx);
    if (width < 1 || 2 * width > N1) return IMPEC_ERR_ARG;
    double val = I / (N2 * dx * dy * width);
    if (mode == 0) { const double tau = 1. / q_freq, x = fmod(t, tau); if (x > tau / 2.) val *= q_high_multiple; }
    else if (mode == 1) { const double vl = val, vh = val * q_high_multiple; val = vl + (t - tstart) * (vh - vl) / (tfinal - tstart); }
    const size_t M = (size_t)N1 * N2;
    for (size_t k = 0; k < M; k++) q[k] = 0.;
    for (size_t k = 0; k < (size_t)N2 * width; k++) { q[k] = val; q[M - 1 - k] = -val; }
    return IMPEC_OK;
}

static inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// Developed synthetic state: c = 1/2 [1 - tanh((r - r0(theta))/w)], r0 = L (0.45 + 0.08 sin 16 theta), w = 4 dx,
// plus 1e-3 (U - 1/2) with U a counter-based splitmix64 stream (decomposition independent), clipped to [0,1];
// the parallel layout adds the mirrored front of the second injector.  Writes the interior of a ghosted
// (N1+2g) x (N2+2g) array.
extern "C" int impec_synth_state_rows(const impec_params *P, unsigned long long seed, int row0, int nrows, double *c_ghosted);
extern "C" int impec_synth_state(const impec_params *P, unsigned long long seed, double *c_ghosted) {
    return impec_synth_state_rows(P, seed, 0, P ? P->N1 : 0, c_ghosted);
}
extern "C" int impec_synth_state_rows(const impec_params *P, unsigned long long seed, int row0, int nrows, double *c_ghosted) {
    if (!P || !c_ghosted || row0 < 0 || row0 + nrows > P->N1) return IMPEC_ERR_ARG;
    const int N1 = P->N1, N2 = P->N2, g = P->g;
    const double dx = (P->x_ub - P->x_lb) / N1, dy = (P->y_ub - P->y_lb) / N2, L = P->x_ub - P->x_lb, w = 4. * dx;
    const int pitch = N2 + 2 * g;
    (void)N1;
    parallel_rows(nrows, [&](int il) {
        const int i = row0 + il;
        for (int j = 0; j < N2; j++) {
            const double x = (i + 0.5) * dx, y = (j + 0.5) * dy;
            double r = sqrt(x * x + y * y), th = atan2(y, x);
            double val = 0.5 * (1. - tanh((r - L * (0.45 + 0.08 * sin(16. * th))) / w));
            if (P->stype == IMPEC_PARALLEL) {
                const double x2 = L - x, y2 = (P->y_ub - P->y_lb) - y;
                r = sqrt(x2 * x2 + y2 * y2); th = atan2(y2, x2);
                const double v2 = 0.5 * (1. - tanh((r - L * (0.45 + 0.08 * sin(16. * th))) / w));
                if (v2 > val) val = v2;
            }
            const unsigned long long h = splitmix64(seed ^ ((unsigned long long)i * (unsigned long long)N2 + (unsigned long long)j));
            val += 1e-3 * ((double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5);
            if (val < 0.) val = 0.;
            if (val > 1.) val = 1.;
            c_ghosted[(size_t)(il + g) * pitch + (j + g)] = val;
        }
    });
    return IMPEC_OK;
}
