/*
 * ts_driver.cc -- TEST INFRASTRUCTURE: runs pieces of the REFERENCE's own time stepper on CPU thread-ranks, so that the oracles
 * of the projection step (oracle/proj_oracle.c) and of the implicit velocity solves (oracle/impl_oracle.c) are pinned against
 * an execution of the reference itself.
 *
 * Linked into oracle/_ref/libsarasts_d3.so with the reference's UNMODIFIED sources, compiled from where they lie under
 * /root/reference: src/solvers/hydro/{hydro,hydro_d3}.cc, src/solvers/scalar/{scalar,scalar_d3}.cc, lib/field/*.cc,
 * lib/boundary/*.cc, lib/force/*.cc, lib/initial/*.cc, lib/io/{parser,tseries}.cc, lib/parallel/{parallel,mpidata}.cc, lib/grid/grid.cc,
 * lib/poisson/*.cc -- against the stand-in Blitz / MPI / yaml-cpp headers of saras_b200/host/compat.  The HDF5-backed classes
 * (writer, reader, probes) are not built: standin_ts/hdf5.h lets their headers parse, and this file defines the members the
 * time stepper references as do-nothing functions (none is called here: no solvePDE, no restart).
 *
 * What runs is the reference's code, called the way hydro_d3::timeAdvance calls it (src/solvers/hydro/hydro_d3.cc):
 *   mode 0  V.divergence(mgRHS, P); mgRHS *= 1.0/dt;                                                    :326-327
 *   mode 1  Pp.syncData(); P += Pp; Pp.gradient(pressureGradient, V); pressureGradient *= dt; V -= pressureGradient;   :347-362
 *   mode 2  imposeUBCs(); imposeVBCs(); imposeWBCs();                                                   :364-366
 *   mode 3  solveVx(); solveVy(); solveVz();                                                            :306-316
 *   mode 4  modes 3, 0, mgSolve, 1, 2 in sequence: the tail of timeAdvance from nseRHS on               :303-366
 *   mode 6  nseRHS = 0; V.computeDiff(nseRHS);  (returned in the nseRHS slots)                          :259-265
 *   mode 7  nSteps x timeAdvance() from the configured initial condition (inputs given replace it)      :254-367
 *   mode 9  nseRHS = 0; V.computeNLin(V, nseRHS);  (returned in the nseRHS slots)                       :268-277
 *   mode 8  scalar_d3::solveT() on the given T and tmpRHS                                     scalar_d3.cc:541-594
 * Fields go in and out as the padded Blitz arrays of every rank.  Ranks are threads (mpi-lite).
 */
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unistd.h>

#define protected public            /* the pieces of hydro_d3 called below are protected members; the sources stay unmodified */
#define private public
#include "hydro.h"
#include "scalar.h"
#undef protected
#undef private

/* ---- do-nothing members of the HDF5-backed classes (see the header comment) ---- */
writer::writer(const grid &mesh, std::vector<field> &wFields) : mesh(mesh), wFields(wFields) {}
void writer::writeTarang(real) {}
void writer::writeSolution(real) {}
void writer::writeRestart(real) {}
writer::~writer() {}
reader::reader(const grid &mesh, std::vector<field> &rFields) : mesh(mesh), rFields(rFields) {}
real reader::readData() { return 0.0; }
reader::~reader() {}
probes::probes(const grid &mesh, std::vector<field> &pFields) : mesh(mesh), pFields(pFields), numFields(0) {}
void probes::probeData(real) {}
probes::~probes() {}

extern "C" {

typedef struct {
    int xInd, yInd, zInd, npX, npY, xPer, yPer, zPer, xGrid, yGrid, zGrid;
    double betaX, betaY, betaZ, Lx, Ly, Lz;
    int vcDepth, vcCount, preSmooth, postSmooth, gsSmooth, resType;
    double mgTolerance;
    int probType;           /* "Problem Type": 1 lid-driven cavity, 3 inflow / outflow, 4 channel, ... (hydro.cc:287-371) */
    double Re, dt, cnTolerance;
    int icType;             /* "Initial Condition": 0 zero, 1 Taylor-Green vortex, 2 / 3 channel profiles (hydro_d3.cc:88-101) */
    int nSteps;             /* mode 7: number of timeAdvance() calls */
    int outputOnly;         /* bit q set: field q of ts_rank_fields is only written (the constructor's / initial condition's value is kept as input) */
    int solverKind;         /* 0 hydro_d3, 1 scalar_d3 (Rayleigh-Benard: Problem Type 5-8) */
    int rbcType, forceType; /* "RBC Type" (non-dimensionalisation, scalar.cc:70-86), "Force" (scalar.cc:131-156: 0 none, 1 random, 2 Coriolis, 3 buoyancy, 4 rotating convection) */
    double Ra, Pr;
} ts_params;

/* per rank: Vx, Vy, Vz, P, Pp, mgRHS, nseRHS.Vx, nseRHS.Vy, nseRHS.Vz, T, tmpRHS (any may be NULL: left as the constructor set it; T, tmpRHS: scalar_d3 only) */
typedef struct { double *f[11]; } ts_rank_fields;

}

namespace {

struct Job { const ts_params *prm; ts_rank_fields *flds; int mode; int status; };

void write_yaml(const ts_params &p, const std::string &dir) {
    std::string cmd = "mkdir -p " + dir + "/input";
    if (system(cmd.c_str()) != 0) fprintf(stderr, "ts_driver: cannot create %s\n", dir.c_str());
    std::ofstream f((dir + "/input/parameters.yaml").c_str());
    char dom[4] = {p.xPer ? 'P' : 'N', p.yPer ? 'P' : 'N', p.zPer ? 'P' : 'N', 0};
    char mesh[4] = {p.xGrid == 2 ? 'D' : 'U', p.yGrid == 2 ? 'D' : 'U', p.zGrid == 2 ? 'D' : 'U', 0};
    f.precision(17);
    f << "\"Program\":\n"
      << "    \"Problem Type\": " << p.probType << "\n    \"Initial Condition\": " << p.icType << "\n    \"Domain Type\": \"" << dom << "\"\n    \"RBC Type\": " << (p.rbcType ? p.rbcType : 1) << "\n"
      << "    \"Reynolds Number\": " << p.Re << "\n    \"Rossby Number\": 12\n    \"Rayleigh Number\": " << (p.Ra > 0 ? p.Ra : 10000.0) << "\n    \"Prandtl Number\": " << (p.Pr > 0 ? p.Pr : 0.71) << "\n"
      << "    \"Taylor Number\": 0.0\n    \"X Length\": " << p.Lx << "\n    \"Y Length\": " << p.Ly << "\n    \"Z Length\": " << p.Lz << "\n"
      << "    \"Heating Plate\": false\n    \"Plate Radius\": 0.25\n    \"Force\": " << p.forceType << "\n"
      << "\"Mesh\":\n    \"Mesh Type\": \"" << mesh << "\"\n    \"X Beta\": " << p.betaX << "\n    \"Y Beta\": " << p.betaY << "\n    \"Z Beta\": " << p.betaZ << "\n"
      << "    \"X Index\": " << p.xInd << "\n    \"Y Index\": " << p.yInd << "\n    \"Z Index\": " << p.zInd << "\n"
      << "\"Parallel\":\n    \"Number of OMP threads\": 1\n    \"X Number of Procs\": " << p.npX << "\n    \"Y Number of Procs\": " << p.npY << "\n"
      << "\"Solver\":\n    \"Differentiation Scheme\": 1\n    \"Integration Scheme\": 1\n    \"Solve Tolerance\": " << p.cnTolerance << "\n    \"Restart Run\": false\n"
      << "    \"Use CFL Condition\": false\n    \"Courant Number\": 0.1\n    \"Time-Step\": " << p.dt << "\n    \"Final Time\": " << p.dt << "\n    \"I/O Count\": 1\n"
      << "    \"Solution Format\": 1\n    \"Solution Write Interval\": 1.0\n    \"Restart Write Interval\": 1.0\n    \"Record Probes\": false\n"
      << "    \"Probe Time Interval\": 0.1\n    \"Probes\": >\n        [5, 5, 5]\n"
      << "\"Multigrid\":\n    \"V-Cycle Depth\": " << p.vcDepth << "\n    \"V-Cycle Count\": " << p.vcCount << "\n    \"Solve Tolerance\": " << p.mgTolerance << "\n"
      << "    \"Smoothing Method\": " << p.gsSmooth << "\n    \"Pre-Smoothing Count\": " << p.preSmooth << "\n    \"Post-Smoothing Count\": " << p.postSmooth << "\n"
      << "    \"Residual Type\": " << p.resType << "\n    \"Print Residual\": false\n";
}

void put(blitz::Array<real, 3> &A, const double *src) { if (src) memcpy(A.dataFirst(), src, A.size() * sizeof(double)); }
void get(const blitz::Array<real, 3> &A, double *dst) { if (dst) memcpy(dst, A.dataFirst(), A.size() * sizeof(double)); }

template <class SOLVER>
void run_solver(Job *job, const grid &gridData, const parser &inputParams, parallel &mpi, blitz::Array<real, 3> *(*extra)(SOLVER &, int), void (*solve_t)(SOLVER &)) {
    const ts_params &p = *job->prm;
    SOLVER H(gridData, inputParams, mpi);        /* P = 1, initial condition, checkPeriodic, initVBC (initTBC), imposeU/V/W(T)BCs */
    H.dt = p.dt;                                 /* solvePDE sets dt = inputParams.tStp (hydro_d3.cc:157); set here as no solvePDE runs */
    ts_rank_fields &F = job->flds[mpi.rank];
    blitz::Array<real, 3> *dst[11] = {&H.V.Vx.F, &H.V.Vy.F, &H.V.Vz.F, &H.P.F.F, &H.Pp.F, &H.mgRHS.F, &H.nseRHS.Vx, &H.nseRHS.Vy, &H.nseRHS.Vz, extra(H, 0), extra(H, 1)};
    for (int q = 0; q < 11; q++) if (dst[q] && !((p.outputOnly >> q) & 1)) put(*dst[q], F.f[q]);
    const int m = job->mode;
    if (m == 6) { H.nseRHS = 0.0; H.V.computeDiff(H.nseRHS); }
    if (m == 9) { H.nseRHS = 0.0; H.V.computeNLin(H.V, H.nseRHS); }  /* the convective term alone, hydro_d3.cc:268-277 */
    if (m == 8) solve_t(H);                                         /* scalar_d3::solveT(), scalar_d3.cc:541-594 */
    if (m == 7) for (int n = 0; n < p.nSteps; n++) H.timeAdvance();
    if (m == 3 || m == 4) { H.solveVx(); H.solveVy(); H.solveVz(); }
    if (m == 0 || m == 4) { H.V.divergence(H.mgRHS, H.P); H.mgRHS *= 1.0 / H.dt; }
    if (m == 4 || m == 5) H.mgSolver.mgSolve(H.Pp, H.mgRHS);      /* mode 5: mgSolve alone (debugging aid) */
    if (m == 1 || m == 4) {
        H.Pp.syncData();
        H.P += H.Pp;
        H.Pp.gradient(H.pressureGradient, H.V);
        H.pressureGradient *= H.dt;
        H.V -= H.pressureGradient;
    }
    if (m == 2 || m == 4) { H.imposeUBCs(); H.imposeVBCs(); H.imposeWBCs(); }
    for (int q = 0; q < 11; q++) if (dst[q]) get(*dst[q], F.f[q]);
}
blitz::Array<real, 3> *no_extra(hydro_d3 &, int) { return nullptr; }
blitz::Array<real, 3> *temperature(scalar_d3 &S, int which) { return which ? &S.tmpRHS.F : &S.T.F.F; }
void no_solve_t(hydro_d3 &) {}
void do_solve_t(scalar_d3 &S) { S.solveT(); }

void rank_main(void *arg) {
    Job *job = (Job *)arg;
    parser inputParams;                          /* reads ./input/parameters.yaml */
    parallel mpi(inputParams);
    grid gridData(inputParams, mpi);
    if (job->prm->solverKind == 1) run_solver<scalar_d3>(job, gridData, inputParams, mpi, temperature, do_solve_t);
    else run_solver<hydro_d3>(job, gridData, inputParams, mpi, no_extra, no_solve_t);
}

}  // namespace

extern "C" {

/* padded extents of the ten fields of one rank, in the order of ts_rank_fields (10 x 3 ints) */
void sarasts_extents(const ts_params *p, int *out) {
    const int n[3] = {(1 << p->xInd) / p->npX + 1, (1 << p->yInd) / p->npY + 1, (1 << p->zInd) + 1};
    const int sc[3] = {n[0] + 2, n[1] + 2, n[2] + 2};
    const int vx[3] = {n[0] + 1, n[1] + 2, n[2] + 2}, vy[3] = {n[0] + 2, n[1] + 1, n[2] + 2}, vz[3] = {n[0] + 2, n[1] + 2, n[2] + 1};
    const int *all[11] = {vx, vy, vz, sc, sc, sc, vx, vy, vz, sc, sc};
    for (int q = 0; q < 11; q++) for (int d = 0; d < 3; d++) out[3 * q + d] = all[q][d];
}

int sarasts_run(const char *workdir, const ts_params *prm, ts_rank_fields *flds, int mode) {
    write_yaml(*prm, workdir);
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd))) return -1;
    if (chdir(workdir) != 0) return -1;
    std::ostringstream captured;
    std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
    Job job = {prm, flds, mode, 0};
    saras_mpi_run(prm->npX * prm->npY, rank_main, &job);
    std::cout.rdbuf(old);
    if (chdir(cwd) != 0) return -1;
    return job.status;
}

}
