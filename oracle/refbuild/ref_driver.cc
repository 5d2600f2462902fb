/*
 * Built a second time with -DSARAS_DROPIN (oracle/_ref/libsarasdropin_*.so): the SAME driver and the reference's own
 * grid.cc / parallel.cc / parser.cc, but "poisson.h" and the solver behind it are this repository's drop-in
 * (saras_b200/host/poisson.{h,cc} over libsarasmg.so).  That is the drop-in claim made executable: the reference's
 * object model, unmodified, constructs `multigrid_d3 mgSolver(gridData, inputParams)` and calls
 * `mgSolver.mgSolve(Pp, mgRHS)` exactly as hydro_d3.cc:65,343 do, and gets the CUDA solver.
 *
 * ref_driver.cc -- runs the REFERENCE's own multigrid solver (lib/poisson/poisson*.cc, lib/grid/grid.cc,
 * lib/parallel/parallel.cc, lib/io/parser.cc compiled unmodified from /root/reference against the stand-in
 * Blitz/MPI/yaml headers of saras_b200/host/compat) and hands inputs/outputs across a C ABI so the Python
 * tests can pin oracle/mg_oracle.c against it.  TEST INFRASTRUCTURE ONLY.
 *
 * One shared object per compile-time mode (PLANAR x TEST_POISSON), see the Makefile.
 * MPI ranks are threads of this process (compat/mpi.h, lock-step sends).
 */
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

/* every standard / compat header the reference headers pull in is included BEFORE the access hack */
#include <blitz/array.h>
#include <mpi.h>
#include <yaml-cpp/yaml.h>
#include <math.h>
#include <sys/time.h>
#include <algorithm>

/* read-only access to the solver's protected coefficient arrays (hx, hxhy, ...); the reference's own
 * translation units are compiled without this, and access specifiers do not change class layout */
#define protected public
#define private public
#include "poisson.h"
#undef protected
#undef private

extern "C" {

/* same layout as mgo_params (oracle/mg_oracle.h) */
typedef struct {
    int dim, test_poisson;
    int xInd, yInd, zInd;
    int npX, npY;
    int xPer, yPer, zPer;
    int xGrid, yGrid, zGrid;
    double betaX, betaY, betaZ;
    double Lx, Ly, Lz;
    int vcDepth, vcCount, preSmooth, postSmooth, gsSmooth, resType;
    double mgTolerance;
    int wall_null;
} ref_params;

}

namespace {

struct Job {
    const ref_params *prm;
    double *const *p_inout;
    const double *const *rhs;
    double *metrics_out;      /* per rank: 6 arrays of level 0 concatenated, each F_d long; may be NULL */
    double *coef_out;         /* 9*(D+1) doubles from rank 0 (hx hy hz hx2 hz2 hzhx hxhy hyhz hxhyhz); may be NULL */
    int mode;                 /* 0 mgSolve, 1 setup only */
    double *res_hist;         /* SARAS_DROPIN: residual history and cycle count taken from the drop-in object (rank 0) */
    int *cycles;
    double *proj_out;         /* mode 2: per rank xi_xStaggr, et_yStaggr, zt_zStaggr (F_d each), xi_xColloc, et_yColloc, zt_zColloc
                                 (F_d - 1 each), then dXi, dEt, dZt -- the first-derivative metrics of the projection step */
};

#ifdef PLANAR
typedef multigrid_d2 mg_t;
#else
typedef multigrid_d3 mg_t;
#endif

void write_yaml(const ref_params &p, const std::string &dir) {
    std::string cmd = "mkdir -p " + dir + "/input";
    if (system(cmd.c_str()) != 0) { fprintf(stderr, "ref_driver: cannot create %s\n", dir.c_str()); }
    std::ofstream f((dir + "/input/parameters.yaml").c_str());
    char dom[4] = {p.xPer ? 'P' : 'N', p.yPer ? 'P' : 'N', p.zPer ? 'P' : 'N', 0};
    char mesh[4] = {p.xGrid == 2 ? 'D' : 'U', p.yGrid == 2 ? 'D' : 'U', p.zGrid == 2 ? 'D' : 'U', 0};
    f.precision(17);
    f << "\"Program\":\n"
      << "    \"Problem Type\": 1\n    \"Initial Condition\": 0\n    \"Domain Type\": \"" << dom << "\"\n    \"RBC Type\": 1\n"
      << "    \"Reynolds Number\": 1000\n    \"Rossby Number\": 12\n    \"Rayleigh Number\": 10000\n    \"Prandtl Number\": 0.71\n"
      << "    \"Taylor Number\": 0.0\n    \"X Length\": " << p.Lx << "\n    \"Y Length\": " << p.Ly << "\n    \"Z Length\": " << p.Lz << "\n"
      << "    \"Heating Plate\": false\n    \"Plate Radius\": 0.25\n    \"Force\": 0\n"
      << "\"Mesh\":\n    \"Mesh Type\": \"" << mesh << "\"\n    \"X Beta\": " << p.betaX << "\n    \"Y Beta\": " << p.betaY << "\n    \"Z Beta\": " << p.betaZ << "\n"
      << "    \"X Index\": " << p.xInd << "\n    \"Y Index\": " << p.yInd << "\n    \"Z Index\": " << p.zInd << "\n"
      << "\"Parallel\":\n    \"Number of OMP threads\": 1\n    \"X Number of Procs\": " << p.npX << "\n    \"Y Number of Procs\": " << p.npY << "\n"
      << "\"Solver\":\n    \"Differentiation Scheme\": 1\n    \"Integration Scheme\": 1\n    \"Solve Tolerance\": 1.0e-6\n    \"Restart Run\": false\n"
      << "    \"Use CFL Condition\": false\n    \"Courant Number\": 0.1\n    \"Time-Step\": 0.001\n    \"Final Time\": 0.001\n    \"I/O Count\": 1\n"
      << "    \"Solution Format\": 1\n    \"Solution Write Interval\": 0.001\n    \"Restart Write Interval\": 0.1\n    \"Record Probes\": false\n"
      << "    \"Probe Time Interval\": 0.1\n    \"Probes\": >\n        [5, 5, 5]\n"
      << "\"Multigrid\":\n    \"V-Cycle Depth\": " << p.vcDepth << "\n    \"V-Cycle Count\": " << p.vcCount << "\n    \"Solve Tolerance\": " << p.mgTolerance << "\n"
      << "    \"Smoothing Method\": " << p.gsSmooth << "\n    \"Pre-Smoothing Count\": " << p.preSmooth << "\n    \"Post-Smoothing Count\": " << p.postSmooth << "\n"
#ifdef SARAS_DROPIN
      << "    \"Residual Type\": " << p.resType << "\n    \"Print Residual\": false\n";      /* history comes from mgSolver.lastResiduals */
#else
      << "    \"Residual Type\": " << p.resType << "\n    \"Print Residual\": true\n";
#endif
}

void rank_main(void *arg) {
    Job *job = (Job *)arg;
    const ref_params &p = *job->prm;
    parser inputParams;                          /* reads ./input/parameters.yaml (parser.cc:72) */
    parallel mpi(inputParams);
    grid gridData(inputParams, mpi);
    mg_t mgSolver(gridData, inputParams);        /* hydro_d3.cc:65: constructed BEFORE checkPeriodic */

    if (p.wall_null) {
        /* hydro::checkPeriodic, hydro.cc:174-203 */
        if (!inputParams.xPer) {
            if (mpi.xRank == 0) mpi.faceRanks(0) = MPI_PROC_NULL;
            if (mpi.xRank == mpi.npX - 1) mpi.faceRanks(1) = MPI_PROC_NULL;
        }
#ifdef PLANAR
        if (mpi.yRank == 0) mpi.faceRanks(2) = MPI_PROC_NULL;
        if (mpi.yRank == mpi.npY - 1) mpi.faceRanks(3) = MPI_PROC_NULL;
#else
        if (!inputParams.yPer) {
            if (mpi.yRank == 0) mpi.faceRanks(2) = MPI_PROC_NULL;
            if (mpi.yRank == mpi.npY - 1) mpi.faceRanks(3) = MPI_PROC_NULL;
        }
#endif
    }

    const int r = mpi.rank;
    if (job->metrics_out) {
        const blitz::Array<real, 1> *m[6] = {&gridData.xixxStaggr, &gridData.xix2Staggr, &gridData.etyyStaggr,
                                              &gridData.ety2Staggr, &gridData.ztzzStaggr, &gridData.ztz2Staggr};
        size_t per_rank = 0;
        for (int q = 0; q < 6; q++) per_rank += m[q]->size();
        double *o = job->metrics_out + (size_t)r * per_rank;
        for (int q = 0; q < 6; q++) for (int i = m[q]->lbound(0); i <= m[q]->ubound(0); i++) *o++ = (*m[q])(i);
    }
    if (job->coef_out && r == 0) {
        double *o = job->coef_out;
        const int D = inputParams.vcDepth;
        for (int l = 0; l <= D; l++) {
            *o++ = mgSolver.hx(l);
#ifdef PLANAR
            *o++ = 0.0;
#else
            *o++ = mgSolver.hy(l);
#endif
            *o++ = mgSolver.hz(l);
#ifdef PLANAR
            *o++ = mgSolver.hx2(l); *o++ = mgSolver.hz2(l); *o++ = mgSolver.hzhx(l); *o++ = 0.0; *o++ = 0.0; *o++ = 0.0;
#else
            *o++ = 0.0; *o++ = 0.0; *o++ = mgSolver.hzhx(l); *o++ = mgSolver.hxhy(l); *o++ = mgSolver.hyhz(l); *o++ = mgSolver.hxhyhz(l);
#endif
        }
    }
    if (job->mode == 2 && job->proj_out) {
        const blitz::Array<real, 1> *m[12] = {&gridData.xi_xStaggr, &gridData.et_yStaggr, &gridData.zt_zStaggr,
                                               &gridData.xi_xColloc, &gridData.et_yColloc, &gridData.zt_zColloc,
                                               &gridData.xix2Staggr, &gridData.ety2Staggr, &gridData.ztz2Staggr,
                                               &gridData.xix2Colloc, &gridData.ety2Colloc, &gridData.ztz2Colloc};
        size_t per_rank = 3;
        for (int q = 0; q < 12; q++) per_rank += m[q]->size();
        double *o = job->proj_out + (size_t)r * per_rank;
        for (int q = 0; q < 6; q++) for (int i = m[q]->lbound(0); i <= m[q]->ubound(0); i++) *o++ = (*m[q])(i);
        *o++ = gridData.dXi; *o++ = gridData.dEt; *o++ = gridData.dZt;
        for (int q = 6; q < 12; q++) for (int i = m[q]->lbound(0); i <= m[q]->ubound(0); i++) *o++ = (*m[q])(i);
    }
    if (job->mode >= 1) return;

    plainsf Pp(gridData), mgRHS(gridData);
    const size_t n = Pp.F.size();
    memcpy(Pp.F.dataFirst(), job->p_inout[r], n * sizeof(double));
    memcpy(mgRHS.F.dataFirst(), job->rhs[r], n * sizeof(double));
    mgSolver.mgSolve(Pp, mgRHS);                 /* hydro_d3.cc:343 */
    memcpy(job->p_inout[r], Pp.F.dataFirst(), n * sizeof(double));
#ifdef SARAS_DROPIN
    if (r == 0) {
        if (job->cycles) *job->cycles = mgSolver.lastCycles;
        if (job->res_hist) for (int i = 0; i < mgSolver.lastCycles; i++) job->res_hist[i] = mgSolver.lastResiduals[i];
    }
#endif
}

}  // namespace

extern "C" {

/* which compile-time mode this library was built in: bit0 PLANAR, bit1 TEST_POISSON */
int sarasref_mode(void) {
    int m = 0;
#ifdef PLANAR
    m |= 1;
#endif
#ifdef TEST_POISSON
    m |= 2;
#endif
#ifdef SARAS_DROPIN
    m |= 4;
#endif
    return m;
}

/* Runs the reference mgSolve on npX*npY thread-ranks.  p_inout[r] / rhs[r]: padded level-0 arrays of rank r.
 * res_hist[>= vcCount] receives the values printed as "Residual after V Cycle i is X" (17 significant digits),
 * max_dev the TEST_POISSON "Maximum absolute deviation" line (6 digits, as printed).  Returns the number of V-cycles
 * the reference ran, or -1 on error.  mode 1: setup only (metrics_out / coef_out); mode 2: setup only, metrics_out receives the
 * first-derivative metrics of grid.cc (layout at Job::proj_out). */
int sarasref_mg_solve(const char *workdir, const ref_params *prm, double *const *p_inout, const double *const *rhs,
                      double *res_hist, double *max_dev, double *metrics_out, double *coef_out, int mode) {
    write_yaml(*prm, workdir);
    char cwd[4096];
    if (!getcwd(cwd, sizeof(cwd))) return -1;
    if (chdir(workdir) != 0) return -1;

    std::ostringstream captured;
    std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
    const std::streamsize oldprec = std::cout.precision(17);

    int dropin_cycles = -1;
    Job job = {prm, p_inout, rhs, mode == 2 ? nullptr : metrics_out, coef_out, mode, res_hist, &dropin_cycles, mode == 2 ? metrics_out : nullptr};
    saras_mpi_run(prm->npX * prm->npY, rank_main, &job);

    std::cout.rdbuf(old);
    std::cout.precision(oldprec);
    if (chdir(cwd) != 0) return -1;

    int cycles = 0;
    std::istringstream is(captured.str());
    std::string line;
    while (std::getline(is, line)) {
        int idx;
        double v;
        if (sscanf(line.c_str(), "Residual after V Cycle %d is %lf", &idx, &v) == 2) {
            if (res_hist) res_hist[idx] = v;
            cycles = idx + 1;
        } else if (sscanf(line.c_str(), "Maximum absolute deviation from analytic solution is: %lf", &v) == 1) {
            if (max_dev) *max_dev = v;
        }
    }
#ifdef SARAS_DROPIN
    if (mode == 0) cycles = dropin_cycles;
#endif
    return cycles;
}

}
