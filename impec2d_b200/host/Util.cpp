// Util.cpp -- test profiles and norms (src/Util.cpp:55-147), source initialisation (src/Impec2DSolver.cpp:241-280),
// plot-file writers in the reference's text format (src/Printer.cpp:98-188: "x y value", blank line per i-row).
#include "Util.h"
#include "DeviceContext.h"
#include <cmath>
#include <fstream>
#include <iostream>
#include <cstdlib>

double Tophat(SimParams const *const sp, double x, double y) {
    const double w = 0.3 * (sp->x_ub - sp->x_lb), h = 0.3 * (sp->y_ub - sp->y_lb);
    const double xc = (sp->x_ub - sp->x_lb) / 2., yc = (sp->y_ub - sp->y_lb) / 2.;
    return (fabs(x - xc) <= w / 2. && fabs(y - yc) <= h / 2.) ? 1. : 0.;
}
double Gaussbell(SimParams const *const sp, double x, double y) {
    const double xc = (sp->x_ub - sp->x_lb) / 2., yc = (sp->y_ub - sp->y_lb) / 2.;
    const double ix = 1. / 16 * sqrt(M_PI) * (-erf(8. * xc - 4. * sp->dx - 8. * x) + erf(8. * xc + 4. * sp->dx - 8. * x));
    const double iy = 1. / 16 * sqrt(M_PI) * (-erf(8. * yc - 4. * sp->dy - 8. * y) + erf(8. * yc + 4. * sp->dy - 8. * y));
    return 1. / (sp->dx * sp->dy) * ix * iy;
}
double Sinewave(SimParams const *const sp, double x, double y) {
    return 1. + 1. / (sp->dx * sp->dy * 4. * M_PI * M_PI) * (cos(2. * M_PI * (x + sp->dx / 2.)) - cos(2. * M_PI * (x - sp->dx / 2.)))
                                                         * (cos(2. * M_PI * (y + sp->dy / 2.)) - cos(2. * M_PI * (y - sp->dy / 2.)));
}
double VerticalSine(SimParams const *const sp, double x, double y) {
    (void)x;
    return 1. + 1. / (sp->dy) * (-1. / (2. * M_PI)) * (cos(2 * M_PI * (y + sp->dy / 2.)) - cos(2 * M_PI * (y - sp->dy / 2.)));
}
double CalcAdvectionL2Norm(SimParams const *const sp, const Array2d &c, double, double (*f)(SimParams const *const, double, double)) {
    double s = 0.;
    for (int i = 0; i < sp->N1; i++) for (int j = 0; j < sp->N2; j++) {
        const double ex = f(sp, sp->x_lb + (i + 0.5) * sp->dx, sp->y_lb + (j + 0.5) * sp->dy);
        s += (ex - c(i, j)) * (ex - c(i, j));
    }
    return sqrt(sp->dx * sp->dy * s);
}
double CalcAdvectionL1Norm(SimParams const *const sp, const Array2d &c, double, double (*f)(SimParams const *const, double, double)) {
    double s = 0.;
    for (int i = 0; i < sp->N1; i++) for (int j = 0; j < sp->N2; j++)
        s += fabs(f(sp, sp->x_lb + (i + 0.5) * sp->dx, sp->y_lb + (j + 0.5) * sp->dy) - c(i, j));
    return s * sp->dx * sp->dy;
}
// InitHorizontalFront (src/Util.cpp:607-639): injection layer at i = 0.., production layer at the far end; pulsing / linear / constant rate
void InitHorizontalFront(SimParams *sp, double *q, double t) {
    const impec_params P = DeviceContext::ToParams(sp);
    const int mode = sp->im == PULSING ? 0 : (sp->im == LINEAR ? 1 : 2);
    const int rc = impec_init_horizontal_front(&P, sp->I, sp->well_width, mode, sp->q_freq, sp->q_high_multiple, sp->tstart, sp->tfinal, t, q);
    if (rc != IMPEC_OK) { std::cerr << "Injection layer cannot be represented on this grid: (qlayer_width < 1)" << std::endl; exit(1); }
    DeviceContext::InvalidateSource(sp);                  // q changed in place: the next solver call uploads it again
}
void InitSource(SimParams *sp, double *q) {
    if (sp->stype == HORIZONTAL) { InitHorizontalFront(sp, q, 0); return; }           // src/Impec2DSolver.cpp:273-274
    if (sp->well_profile == GAUSSIAN) { for (int k = 0; k < sp->M; k++) q[k] = 0.; return; }   // no-op upstream (:253-254)
    const impec_params P = DeviceContext::ToParams(sp);
    const int rc = impec_init_source(&P, sp->I, sp->well_rc, sp->well_profile == ONECELL ? 2 : 0, q);
    if (rc != IMPEC_OK) { std::cerr << "SOURCE-TERM: unknown or unsupported source term!" << std::endl; exit(1); }
}
void PrintDToPlotFile(const std::string &file, const Array2d &a, double x_lb, double y_lb, double dx, double dy, int bounds) {
    std::ofstream f(file.c_str());
    if (!f.is_open()) { std::cerr << "Unable to open file " << file << std::endl; return; }
    const int N1 = a.extent(0) - 2 * bounds, N2 = a.extent(1) - 2 * bounds;
    for (int i = 0; i < N1; i++) {
        for (int j = 0; j < N2; j++) f << x_lb + (i + 0.5) * dx << " " << y_lb + (j + 0.5) * dy << " " << a(i, j) << "\n";
        f << "\n";
    }
}
void PrintVelToPlotFile(const std::string &ufile, const std::string &vfile, const Array3d &v, double x_lb, double y_lb, double dx, double dy, int, int N1, int N2) {
    std::ofstream fu(ufile.c_str()), fv(vfile.c_str());
    if (!fu.is_open() || !fv.is_open()) { std::cerr << "Unable to open velocity files" << std::endl; return; }
    for (int i = 0; i <= N1; i++) { for (int j = 0; j < N2; j++) fu << x_lb + i * dx << " " << y_lb + (j + 0.5) * dy << " " << v(i, j, 0) << "\n"; fu << "\n"; }
    for (int i = 0; i < N1; i++) { for (int j = 0; j <= N2; j++) fv << x_lb + (i + 0.5) * dx << " " << y_lb + j * dy << " " << v(i, j, 1) << "\n"; fv << "\n"; }
}
