// Util.h -- host-side helpers with the reference's names (src/Util.h): test profiles, norms, source initialisation.
#ifndef UTIL_H_
#define UTIL_H_
#include "SimParams.h"
#include "GenSimDefs.h"
double Tophat(SimParams const *const sp, double x, double y);
double Gaussbell(SimParams const *const sp, double x, double y);
double VerticalSine(SimParams const *const sp, double x, double y);
double Sinewave(SimParams const *const sp, double x, double y);
double CalcAdvectionL2Norm(SimParams const *const sp, const Array2d &c, double t, double (*fptr)(SimParams const *const, double, double));
double CalcAdvectionL1Norm(SimParams const *const sp, const Array2d &c, double t, double (*fptr)(SimParams const *const, double, double));
void InitHorizontalFront(SimParams *sp, double *q, double t);   // src/Util.h:28
void InitSource(SimParams *sp, double *q);            // Impec2DSolver::InitializeSource (circle / one_cell wells)
void PrintDToPlotFile(const std::string &file, const Array2d &a, double x_lb, double y_lb, double dx, double dy, int bounds);
void PrintVelToPlotFile(const std::string &ufile, const std::string &vfile, const Array3d &v, double x_lb, double y_lb, double dx, double dy, int bounds, int N1, int N2);
#endif
