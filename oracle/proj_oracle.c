/*
 * proj_oracle.c -- CPU restatement of the two field operations that bracket poisson::mgSolve in SARAS's projection step
 * (SURVEY.md 8f-1).  TEST INFRASTRUCTURE ONLY: called from tests/ (and the smoke test), never from the product.
 *
 * Restated from (paths relative to the SARAS tree):
 *   vfield::divergence                  lib/field/vfield.cc:299-320        + `mgRHS *= 1.0/dt`     src/solvers/hydro/hydro_d3.cc:327
 *   mpidata::syncData (one rank)        lib/parallel/mpidata.cc:82-261, 278-306; faceRanks / edgeRanks: lib/parallel/parallel.cc:101-131,
 *                                       hydro::checkPeriodic src/solvers/hydro/hydro.cc:174-203
 *   sfield::operator+=(plainsf&)        lib/field/sfield.cc:197-201
 *   plainsf::gradient                   lib/field/plainsf.h:92-98          + `pressureGradient *= dt; V -= pressureGradient`  hydro_d3.cc:360-362
 *   collocated metric xi_xColloc        lib/grid/grid.cc:370-422 (uniform), 442-519 (tanh)
 *
 * PINNED against an execution of the reference itself.  oracle/_ref/libsarasts_d3.so is the reference's UNMODIFIED
 * src/solvers/hydro, lib/field, lib/boundary, lib/parallel/mpidata.cc ... compiled by oracle/refbuild/Makefile (stand-in Blitz /
 * MPI / yaml-cpp headers; oracle/refbuild/ts_driver.cc calls the statements cited above the way hydro_d3::timeAdvance does, on
 * CPU thread-ranks).  tests/test_oracle_vs_timestepper.py requires every function of this file to reproduce it BIT FOR BIT:
 * divergence, syncData + correction on 1, 2x1, 2x2, 4x2 and 1x2 ranks (uniform / tanh, periodic / walls), and the whole tail
 * of timeAdvance chained with the multigrid and the implicit-solve oracles.  The metric restatements are pinned against the
 * reference's grid object as well (tests/test_proj_oracle.py).
 *
 * Array layouts are the reference's Blitz arrays, base index -1 in every direction, z fastest (lib/field/field.cc:66-105):
 *   staggered scalar (P, Pp, mgRHS)  (nx+2, ny+2, nz+2);  Vx (nx+1, ny+2, nz+2);  Vy (nx+2, ny+1, nz+2);  Vz (nx+2, ny+2, nz+1)
 * with nx, ny, nz the staggered core sizes (2^M + 1).
 */
#include <math.h>
#include <stddef.h>

#define SC(i, j, k) ((((size_t)((i) + 1)) * (ny + 2) + ((j) + 1)) * (nz + 2) + ((k) + 1))
#define VX(i, j, k) ((((size_t)((i) + 1)) * (ny + 2) + ((j) + 1)) * (nz + 2) + ((k) + 1))
#define VY(i, j, k) ((((size_t)((i) + 1)) * (ny + 1) + ((j) + 1)) * (nz + 2) + ((k) + 1))
#define VZ(i, j, k) ((((size_t)((i) + 1)) * (ny + 2) + ((j) + 1)) * (nz + 1) + ((k) + 1))

/* grid.cc:404-421 / 442-519: xi_x at the COLLOCATED points of one direction.  out has nCoreStag + 1 entries, out[0] is index -1.
 * coreStart = first global index of the rank, dxi = 1 / (global intervals), gridType 2 = tanh-stretched (core points only:
 * the pads keep the uniform value, exactly like the staggered arrays). */
void projo_colloc_metric(int nCoreStag, int coreStart, int gridType, double beta, double len, double dxi, double *out) {
    const int nColl = nCoreStag - 1;
    for (int i = -1; i <= nColl; i++) out[i + 1] = 1.0 / len;
    if (gridType == 2) {
        for (int i = 0; i < nColl; i++) {
            const double xi0 = (double)(coreStart + i) * dxi, xi1 = (double)(coreStart + i + 1) * dxi;
            const double x = len * (1.0 - tanh(beta * (1.0 - (xi0 + xi1))) / tanh(beta)) / 2.0;
            out[i + 1] = tanh(beta) / (beta * len * (1.0 - pow((1.0 - 2.0 * x / len) * tanh(beta), 2)));
        }
    }
}

/* grid.cc:404-406 / 458-466: xi_x at the STAGGERED points.  out has nCoreStag + 2 entries, out[0] is index -1. */
void projo_staggr_metric(int nCoreStag, int coreStart, int gridType, double beta, double len, double dxi, double *out) {
    for (int i = -1; i <= nCoreStag; i++) out[i + 1] = 1.0 / len;
    if (gridType == 2) {
        for (int i = 0; i < nCoreStag; i++) {
            const double xi = (double)(coreStart + i) * dxi;
            const double x = len * (1.0 - tanh(beta * (1.0 - 2.0 * xi)) / tanh(beta)) / 2.0;
            out[i + 1] = tanh(beta) / (beta * len * (1.0 - pow((1.0 - 2.0 * x / len) * tanh(beta), 2)));
        }
    }
}

/* vfield.cc:308-319: divV = 0; divV(core) = xi_x(i)*(Vx - Vx(i-1))/dXi + et_y(j)*(Vy - Vy(j-1))/dEt + zt_z(k)*(Vz - Vz(k-1))/dZt;
 * hydro_d3.cc:327: mgRHS *= 1.0/dt (whole array).  xs/ys/zs: staggered metrics, entry [i+1] is index i. */
void projo_divergence(int nx, int ny, int nz, const double *vx, const double *vy, const double *vz,
                      const double *xs, const double *ys, const double *zs, double dXi, double dEt, double dZt,
                      double inv_dt, double *rhs) {
    const size_t n = (size_t)(nx + 2) * (ny + 2) * (nz + 2);
    for (size_t q = 0; q < n; q++) rhs[q] = 0.0;
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++)
        rhs[SC(i, j, k)] = xs[i + 1] * (vx[VX(i, j, k)] - vx[VX(i - 1, j, k)]) / dXi +
                           ys[j + 1] * (vy[VY(i, j, k)] - vy[VY(i, j - 1, k)]) / dEt +
                           zs[k + 1] * (vz[VZ(i, j, k)] - vz[VZ(i, j, k - 1)]) / dZt;
    for (size_t q = 0; q < n; q++) rhs[q] *= inv_dt;
}

/* mpidata::syncData of a staggered scalar on ONE rank.  Faces (core rows only, mpidata.cc:105-178) go to faceRanks: the rank
 * itself along a periodic direction, MPI_PROC_NULL otherwise (hydro.cc:183-201).  Edges (mpidata.cc:183-259) go to edgeRanks,
 * which checkPeriodic never nulls: with one rank they are the rank itself whatever the domain type (parallel.cc:124-129).
 * Staggered fields share the boundary point, so plane 1 / n-2 is sent (mpidata.cc:112, 131). */
void projo_sync_one_rank(int nx, int ny, int nz, int xper, int yper, double *A) {
    if (xper)
        for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++) {
            const double left = A[SC(nx - 2, j, k)], right = A[SC(1, j, k)];        /* the two messages are sent before they land */
            A[SC(-1, j, k)] = left; A[SC(nx, j, k)] = right;
        }
    if (yper)
        for (int i = 0; i < nx; i++) for (int k = 0; k < nz; k++) {
            const double front = A[SC(i, ny - 2, k)], back = A[SC(i, 1, k)];
            A[SC(i, -1, k)] = front; A[SC(i, ny, k)] = back;
        }
    for (int k = 0; k < nz; k++) {
        const double c00 = A[SC(nx - 2, ny - 2, k)], c01 = A[SC(nx - 2, 1, k)], c10 = A[SC(1, ny - 2, k)], c11 = A[SC(1, 1, k)];
        A[SC(-1, -1, k)] = c00; A[SC(-1, ny, k)] = c01; A[SC(nx, -1, k)] = c10; A[SC(nx, ny, k)] = c11;
    }
}

/* hydro_d3.cc:347-362: Pp.syncData(); P += Pp; Pp.gradient(pressureGradient, V); pressureGradient *= dt; V -= pressureGradient.
 * xc/yc/zc: collocated metrics, entry [i+1] is index i.  pp is modified by syncData, P and V in place. */
void projo_correct(int nx, int ny, int nz, int xper, int yper, double *pp, double *P, double *vx, double *vy, double *vz,
                   const double *xc, const double *yc, const double *zc, double dXi, double dEt, double dZt, double dt) {
    projo_sync_one_rank(nx, ny, nz, xper, yper, pp);
    const size_t n = (size_t)(nx + 2) * (ny + 2) * (nz + 2);
    for (size_t q = 0; q < n; q++) P[q] += pp[q];
    for (int i = 0; i <= nx - 2; i++) for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++) {
        double g = xc[i + 1] * (pp[SC(i + 1, j, k)] - pp[SC(i, j, k)]) / dXi;
        g *= dt;
        vx[VX(i, j, k)] -= g;
    }
    for (int i = 0; i < nx; i++) for (int j = 0; j <= ny - 2; j++) for (int k = 0; k < nz; k++) {
        double g = yc[j + 1] * (pp[SC(i, j + 1, k)] - pp[SC(i, j, k)]) / dEt;
        g *= dt;
        vy[VY(i, j, k)] -= g;
    }
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int k = 0; k <= nz - 2; k++) {
        double g = zc[k + 1] * (pp[SC(i, j, k + 1)] - pp[SC(i, j, k)]) / dZt;
        g *= dt;
        vz[VZ(i, j, k)] -= g;
    }
}

/* mpidata::syncData of a staggered scalar on npX x npY ranks (mpidata.cc:278-306), all ranks emulated in one process:
 * A[r] is rank r's array (rank = yRank * npX + xRank, parallel.cc:89-92).  Faces go to faceRanks -- the periodic wrap,
 * nulled at the walls of a non-periodic direction by hydro::checkPeriodic (hydro.cc:174-203) -- and carry core rows only;
 * edges go to edgeRanks, the periodic wrap whatever the domain type (parallel.cc:124-129).  Every message is read from core
 * points and written to pads, so the order of the eight messages does not matter. */
static int projo_find_rank(int xr, int yr, int npX, int npY) { return ((yr + npY) % npY) * npX + ((xr + npX) % npX); }
void projo_sync_multi(int npX, int npY, int nx, int ny, int nz, int xper, int yper, double *const *A) {
    for (int r = 0; r < npX * npY; r++) {
        const int xr = r % npX, yr = r / npX;
        double *D = A[r];
        if (xper || xr > 0) {               /* left pad <- left neighbour's plane nx-2 */
            const double *S = A[projo_find_rank(xr - 1, yr, npX, npY)];
            for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++) D[SC(-1, j, k)] = S[SC(nx - 2, j, k)];
        }
        if (xper || xr < npX - 1) {
            const double *S = A[projo_find_rank(xr + 1, yr, npX, npY)];
            for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++) D[SC(nx, j, k)] = S[SC(1, j, k)];
        }
        if (yper || yr > 0) {
            const double *S = A[projo_find_rank(xr, yr - 1, npX, npY)];
            for (int i = 0; i < nx; i++) for (int k = 0; k < nz; k++) D[SC(i, -1, k)] = S[SC(i, ny - 2, k)];
        }
        if (yper || yr < npY - 1) {
            const double *S = A[projo_find_rank(xr, yr + 1, npX, npY)];
            for (int i = 0; i < nx; i++) for (int k = 0; k < nz; k++) D[SC(i, ny, k)] = S[SC(i, 1, k)];
        }
        for (int xh = 0; xh < 2; xh++) for (int yh = 0; yh < 2; yh++) {
            const double *S = A[projo_find_rank(xr + (xh ? 1 : -1), yr + (yh ? 1 : -1), npX, npY)];
            for (int k = 0; k < nz; k++) D[SC(xh ? nx : -1, yh ? ny : -1, k)] = S[SC(xh ? 1 : nx - 2, yh ? 1 : ny - 2, k)];
        }
    }
}

/* the correction step of one rank AFTER its Pp has been synchronised (projo_sync_multi): P += Pp; V -= dt grad(Pp) */
void projo_correct_nosync(int nx, int ny, int nz, const double *pp, double *P, double *vx, double *vy, double *vz,
                          const double *xc, const double *yc, const double *zc, double dXi, double dEt, double dZt, double dt) {
    const size_t n = (size_t)(nx + 2) * (ny + 2) * (nz + 2);
    for (size_t q = 0; q < n; q++) P[q] += pp[q];
    for (int i = 0; i <= nx - 2; i++) for (int j = 0; j < ny; j++) for (int k = 0; k < nz; k++) {
        double g = xc[i + 1] * (pp[SC(i + 1, j, k)] - pp[SC(i, j, k)]) / dXi;
        g *= dt;
        vx[VX(i, j, k)] -= g;
    }
    for (int i = 0; i < nx; i++) for (int j = 0; j <= ny - 2; j++) for (int k = 0; k < nz; k++) {
        double g = yc[j + 1] * (pp[SC(i, j + 1, k)] - pp[SC(i, j, k)]) / dEt;
        g *= dt;
        vy[VY(i, j, k)] -= g;
    }
    for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int k = 0; k <= nz - 2; k++) {
        double g = zc[k + 1] * (pp[SC(i, j, k + 1)] - pp[SC(i, j, k)]) / dZt;
        g *= dt;
        vz[VZ(i, j, k)] -= g;
    }
}
