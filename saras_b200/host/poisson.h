/*
 * poisson.h -- DROP-IN replacement for SARAS's lib/poisson/poisson.h (classes poisson, multigrid_d2, multigrid_d3).
 *
 * Same class names, same constructors `(const grid &mesh, const parser &solParam)`, same public surface
 * (mgSolve, testProlong, testTransfer, testPeriodic, virtual destructor; reference poisson.h:125-135, 156-219), so
 * hydro_d2/d3 and scalar_d2/d3 (`multigrid_d3 mgSolver;` hydro.h:193, `mgSolver.mgSolve(Pp, mgRHS);` hydro_d3.cc:343)
 * compile and call it unchanged.  The solver itself runs on the GPU behind the C ABI of include/sarasmg.h: this
 * class only (1) evaluates the reference's host-side setup expressions -- level sizes, the pow() coefficient
 * products, the strided copies of the grid metrics, the TEST_POISSON wall planes (reference poisson.cc:326-496,
 * poisson_d3.cc:513-572, poisson_d2.cc:367-400) -- and (2) forwards mgSolve.  The device hierarchy is created
 * lazily at the first mgSolve because the callers edit mesh.rankData.faceRanks after constructing the solver
 * (hydro_d3.cc:65 vs :108).
 *
 * Compile-time modes keep their meaning: -DPLANAR (grid is 2-D; use multigrid_d2), -DTEST_POISSON (Dirichlet
 * r^2/6 test problem, no mean removal, no early exit, analytic check printed), -DTIME_RUN (compute/transfer
 * split of smooth() printed by the destructor).  -DREAL_SINGLE is rejected: the device path is fp64.
 */
#ifndef POISSON_H
#define POISSON_H

#include <blitz/array.h>
#include <math.h>
#include <vector>

#include "plainsf.h"
#include "grid.h"

static_assert(sizeof(real) == sizeof(double), "the B200 multigrid path is fp64 only: build SARAS with -DREAL_DOUBLE");

struct sarasmg_solver_s;

class poisson {
   protected:
    const grid &mesh;
    const parser &inputParams;

    int dim;                                        /* 2: multigrid_d2, 3: multigrid_d3 */
    int maxCount;
    blitz::TinyVector<int, 3> localSizeIndex;
    blitz::Array<int, 1> xEnd, yEnd, zEnd;          /* per level, as in the reference (poisson.h:77) */
    blitz::Array<int, 1> strideValues;
    blitz::Array<real, 1> hx, hy, hz, hx2, hz2, hzhx, hxhy, hyhz, hxhyhz;
    std::vector<double> xixx, xix2, etyy, ety2, ztzz, ztz2;    /* per level padded arrays, concatenated */
    std::vector<double> xWall, yWall, zWall;

    sarasmg_solver_s *device;                       /* created at the first mgSolve */
    std::vector<const void *> pinnedFields;         /* caller field storage page-locked by mgSolve, released in ~poisson */
    void pinField(const void *ptr, unsigned long long bytes);

    void setLocalSizeIndex();
    void setStagBounds();
    void setCoefficients();
    void copyStaggrDerivs();
    void initDirichlet();
    void setup(int dimension);
    void ensureDevice();

   public:
    poisson(const grid &mesh, const parser &solParam);

    void mgSolve(plainsf &inFn, const plainsf &rhs);

    virtual real testProlong();
    virtual real testTransfer();
    virtual real testPeriodic();

    /* extras (not in the reference): results of the last mgSolve and the native handle for instrumentation */
    int lastCycles;
    std::vector<double> lastResiduals;
    std::vector<int> lastCoarseIters;
    sarasmg_solver_s *nativeHandle() { ensureDevice(); return device; }

    virtual ~poisson();
};

class multigrid_d2 : public poisson {
   public:
    multigrid_d2(const grid &mesh, const parser &solParam);
    ~multigrid_d2() {}
};

class multigrid_d3 : public poisson {
   public:
    multigrid_d3(const grid &mesh, const parser &solParam);
    ~multigrid_d3() {}
};

#endif
