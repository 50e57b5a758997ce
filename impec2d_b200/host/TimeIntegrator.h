// TimeIntegrator / TRIntegrator -- source-term integration, same surface as src/TimeIntegrator.h:6-16 and
// src/TRIntegrator.h.  Integrate() runs the trapezoidal/forward update of src/TRIntegrator.cpp:49-71 on the GPU
// over the compact well list (impec_source_split).
#ifndef TIMEINTEGRATOR_H_
#define TIMEINTEGRATOR_H_
#include "SimParams.h"
#include "GenSimDefs.h"
class TimeIntegrator {
protected:
    SimParams *sp;
    double total_mass_outflux;
public:
    TimeIntegrator(SimParams *sp) : sp(sp), total_mass_outflux(0) {}
    virtual ~TimeIntegrator() {}
    double GetTotalMassOutflux() { return total_mass_outflux; }
    virtual int Integrate(Array2d &c, double const *const q, const double dt) = 0;
    int IntegrateSubcycle(Array2d &c, double const *const q, const double dt) { return Integrate(c, q, dt); }
};
class TRIntegrator : public TimeIntegrator {
public:
    TRIntegrator(SimParams *sp) : TimeIntegrator(sp) {}
    virtual ~TRIntegrator() {}
    virtual int Integrate(Array2d &c, double const *const q, const double dt);
};
#endif
