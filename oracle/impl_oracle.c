/*
 * impl_oracle.c -- CPU restatement of SARAS's implicit (Crank-Nicolson) velocity solves, the Jacobi iterations
 * hydro_d3::solveVx / solveVy / solveVz.  TEST INFRASTRUCTURE ONLY: called from tests/, never from the product.
 *
 * Restated from (paths relative to the SARAS tree):
 *   hydro_d3::solveVx / solveVy / solveVz      src/solvers/hydro/hydro_d3.cc:370-533
 *   hydro::setCoefficients, maxIterations      src/solvers/hydro/hydro.cc:223-241, 74
 *   hydro::imposeUBCs / VBCs / WBCs            src/solvers/hydro/hydro.cc:386-445 (order: left, right, front, back, TOP, BOTTOM)
 *   hydro::initVBC (which class on which wall) src/solvers/hydro/hydro.cc:287-371
 *   boundary classes                            lib/boundary/{dirichletCC,dirichletFC,neumannCC,neumannFC,periodicCC,periodicFC}.cc:71-84
 *   field bulk / wall slices                    lib/field/field.cc:180-226, 249-290
 *   mpidata::syncData on one rank               lib/parallel/mpidata.cc:82-306 (collocated directions send plane 0 / n-1, staggered 1 / n-2)
 *   plainvf::vxMax                              lib/field/plainvf.h:97-105 (max over the WHOLE array; its non-bulk entries stay 0)
 *
 * PINNED against an execution of the reference itself: tests/test_oracle_vs_timestepper.py runs the reference's own
 * hydro_d3::solveVx / solveVy / solveVz and hydro::imposeUBCs / VBCs / WBCs (oracle/_ref/libsarasts_d3.so: the unmodified
 * sources compiled by oracle/refbuild/Makefile, driven by oracle/refbuild/ts_driver.cc) and requires this file to reproduce
 * the resulting velocity fields BIT FOR BIT -- lid-driven cavity, channel, fully periodic and inflow / outflow boundary sets,
 * uniform and tanh meshes, alone and chained with the projection step.  The squared metrics are pinned against the reference's
 * grid object (tests/test_impl_oracle.py).  One rank (implo_solve), npX x npY ranks (implo_solve_multi), and the scalar solver's
 * temperature solve (implo_solve_T, one rank).
 *
 * A velocity component is a Blitz array with base index -1, z fastest: extent of a direction = core size + 2, where the core
 * has n points if the component is staggered (cell-centred) there and n - 1 if it is collocated (face-centred): Vx is
 * collocated in x only, Vy in y, Vz in z (lib/field/vfield.cc:60-66).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { IMPL_BC_NONE = 0, IMPL_BC_DIRICHLET_CC, IMPL_BC_DIRICHLET_FC, IMPL_BC_NEUMANN, IMPL_BC_PERIODIC_CC, IMPL_BC_PERIODIC_FC };

typedef struct {
    int n[3];               /* staggered core sizes of the rank (2^M + 1) */
    int per[3];             /* xPer, yPer, zPer */
    double dt, Re, tol;     /* inputParams.dt (or the CFL step), Re, cnTolerance */
    double dXi, dEt, dZt;
    const double *m2c[3];   /* xix2Colloc, ety2Colloc, ztz2Colloc: n - 1 + 2 entries, [0] is index -1 */
    const double *m2s[3];   /* xix2Staggr, ety2Staggr, ztz2Staggr: n + 2 entries */
    int bc_type[6];         /* of THIS component: left, right, front, back, bottom, top */
    double bc_value[6];
} implo_params;

typedef struct { int cs[3], ext[3], lb[3], ub[3], stag[3]; } implo_geom;

static void implo_geometry(const implo_params *p, int comp, implo_geom *g) {
    for (int d = 0; d < 3; d++) {
        g->stag[d] = (d != comp);
        g->cs[d] = g->stag[d] ? p->n[d] : p->n[d] - 1;
        g->ext[d] = g->cs[d] + 2;
        const int cut = (g->stag[d] && !p->per[d]) ? 1 : 0;            /* field.cc:206-219, one rank */
        g->lb[d] = cut; g->ub[d] = g->cs[d] - 1 - cut;
    }
}
#define AT(g, i, j, k) ((((size_t)((i) + 1)) * (g)->ext[1] + ((j) + 1)) * (g)->ext[2] + ((k) + 1))

/* plane `dst` of axis `ax` <- a + b * plane `src` over the whole extent of the other two axes (b == 0: constant; copy: a = 0, b = 1 exactly) */
static void implo_plane(double *F, const implo_geom *g, int ax, int dst, int src, int mode, double v) {
    int lo[3] = {-1, -1, -1}, hi[3] = {g->ext[0] - 2, g->ext[1] - 2, g->ext[2] - 2};
    lo[ax] = hi[ax] = dst;
    for (int i = lo[0]; i <= hi[0]; i++) for (int j = lo[1]; j <= hi[1]; j++) for (int k = lo[2]; k <= hi[2]; k++) {
        int s[3] = {i, j, k};
        s[ax] = src;
        const double f = F[AT(g, s[0], s[1], s[2])];
        double r;
        if (mode == IMPL_BC_DIRICHLET_CC) r = v;
        else if (mode == IMPL_BC_DIRICHLET_FC) r = 2.0 * v - f;
        else r = f;
        F[AT(g, i, j, k)] = r;
    }
}

/* imposeUBCs / VBCs / WBCs on one rank: syncData, then the walls in the order left, right, front, back, top, bottom */
void implo_impose_bcs(const implo_params *p, int comp, double *F) {
    implo_geom G, *g = &G;
    implo_geometry(p, comp, g);
    /* syncData: faces with faceRanks (self along a periodic axis, else MPI_PROC_NULL), core rows only; then the four edges
     * with edgeRanks (always self on one rank) */
    for (int ax = 0; ax < 2; ax++) {
        if (!p->per[ax]) continue;
        const int o = g->stag[ax] ? 1 : 0, cs = g->cs[ax], a2 = 1 - ax;
        for (int q = 0; q < g->cs[a2]; q++) for (int k = 0; k < g->cs[2]; k++) {
            int lo[3], hi[3], sl[3], sh[3];
            lo[ax] = -1; hi[ax] = cs; sl[ax] = cs - 1 - o; sh[ax] = o;
            lo[a2] = hi[a2] = sl[a2] = sh[a2] = q;
            lo[2] = hi[2] = sl[2] = sh[2] = k;
            const double left = F[AT(g, sl[0], sl[1], sl[2])], right = F[AT(g, sh[0], sh[1], sh[2])];
            F[AT(g, lo[0], lo[1], lo[2])] = left; F[AT(g, hi[0], hi[1], hi[2])] = right;
        }
    }
    {
        const int ox = g->stag[0] ? 1 : 0, oy = g->stag[1] ? 1 : 0, cx = g->cs[0], cy = g->cs[1];
        for (int k = 0; k < g->cs[2]; k++) {
            const double c00 = F[AT(g, cx - 1 - ox, cy - 1 - oy, k)], c01 = F[AT(g, cx - 1 - ox, oy, k)];
            const double c10 = F[AT(g, ox, cy - 1 - oy, k)], c11 = F[AT(g, ox, oy, k)];
            F[AT(g, -1, -1, k)] = c00; F[AT(g, -1, cy, k)] = c01; F[AT(g, cx, -1, k)] = c10; F[AT(g, cx, cy, k)] = c11;
        }
    }
    const int order[6] = {0, 1, 2, 3, 5, 4};
    for (int q = 0; q < 6; q++) {
        const int w = order[q], ax = w / 2, hi = w & 1;
        const int t = p->bc_type[w];
        if (t == IMPL_BC_NONE) continue;
        const int wall_lo = g->lb[ax] - 1, wall_hi = g->ub[ax] + 1;     /* field.cc:270-285 */
        const int dst = hi ? wall_hi : wall_lo;
        int src = dst + (hi ? -1 : 1);                                  /* shiftVal (boundary.cc:55-84) */
        if (t == IMPL_BC_PERIODIC_CC) src = hi ? wall_lo + 2 : wall_hi - 2;
        if (t == IMPL_BC_PERIODIC_FC) src = hi ? wall_lo + 1 : wall_hi - 1;
        implo_plane(F, g, ax, dst, src, t, p->bc_value[w]);
    }
}

/* solveVx / solveVy / solveVz: V (in/out) and rhs = nseRHS component, both in the component's padded layout.  Returns the number
 * of Jacobi sweeps performed (iterCount + 1 of the reference), or -1 when maxIterations is exceeded; *max_err = last maxError. */
int implo_solve(const implo_params *p, int comp, double *V, const double *rhs, double *max_err) {
    implo_geom G, *g = &G;
    implo_geometry(p, comp, g);
    const size_t tot = (size_t)g->ext[0] * g->ext[1] * g->ext[2];
    double *guess = (double *)calloc(tot, sizeof(double));              /* guessedVelocity: zero outside the bulk, always */
    const double hx = p->dXi, hy = p->dEt, hz = p->dZt;
    const double hz2hx2 = pow(p->dZt, 2.0) * pow(p->dXi, 2.0), hx2hy2 = pow(p->dXi, 2.0) * pow(p->dEt, 2.0), hy2hz2 = pow(p->dEt, 2.0) * pow(p->dZt, 2.0);
    const double hx2hy2hz2 = pow(p->dXi, 2.0) * pow(p->dEt, 2.0) * pow(p->dZt, 2.0);
    const double inverseRe = 1.0 / p->Re, dt = p->dt, Re = p->Re;
    const double *mx = (g->stag[0] ? p->m2s[0] : p->m2c[0]) + 1, *my = (g->stag[1] ? p->m2s[1] : p->m2c[1]) + 1, *mz = (g->stag[2] ? p->m2s[2] : p->m2c[2]) + 1;
    const long long maxIterations = (long long)(p->n[0] - 1) * (p->n[1] - 1) * (p->n[2] - 1);      /* collocCoreSize product, hydro.cc:74 */
    int iterCount = 0;
    double maxError = 0.0;
    for (;;) {
        for (int i = g->lb[0]; i <= g->ub[0]; i++) for (int j = g->lb[1]; j <= g->ub[1]; j++) for (int k = g->lb[2]; k <= g->ub[2]; k++)
            guess[AT(g, i, j, k)] = ((hy2hz2 * mx[i] * (V[AT(g, i + 1, j, k)] + V[AT(g, i - 1, j, k)]) +
                                      hz2hx2 * my[j] * (V[AT(g, i, j + 1, k)] + V[AT(g, i, j - 1, k)]) +
                                      hx2hy2 * mz[k] * (V[AT(g, i, j, k + 1)] + V[AT(g, i, j, k - 1)])) *
                                     dt / (hx2hy2hz2 * 2.0 * Re) + rhs[AT(g, i, j, k)]) /
                                    (1.0 + dt * ((hy2hz2 * mx[i] + hz2hx2 * my[j] + hx2hy2 * mz[k])) / (Re * hx2hy2hz2));
        memcpy(V, guess, tot * sizeof(double));                         /* V.Vx.F = guessedVelocity.Vx */
        implo_impose_bcs(p, comp, V);
        maxError = 0.0;                                                 /* non-bulk entries of velocityLaplacian are 0 */
        for (int i = g->lb[0]; i <= g->ub[0]; i++) for (int j = g->lb[1]; j <= g->ub[1]; j++) for (int k = g->lb[2]; k <= g->ub[2]; k++) {
            const double c = V[AT(g, i, j, k)];
            const double lap = c - (mx[i] * (V[AT(g, i + 1, j, k)] - 2.0 * c + V[AT(g, i - 1, j, k)]) / (hx * hx) +
                                    my[j] * (V[AT(g, i, j + 1, k)] - 2.0 * c + V[AT(g, i, j - 1, k)]) / (hy * hy) +
                                    mz[k] * (V[AT(g, i, j, k + 1)] - 2.0 * c + V[AT(g, i, j, k - 1)]) / (hz * hz)) *
                                       0.5 * dt * inverseRe;
            const double e = fabs(lap - rhs[AT(g, i, j, k)]);
            if (e > maxError) maxError = e;
        }
        if (maxError < p->tol) break;
        iterCount += 1;
        if (iterCount > maxIterations) { free(guess); if (max_err) *max_err = maxError; return -1; }
    }
    free(guess);
    if (max_err) *max_err = maxError;
    return iterCount + 1;
}

/* second-derivative metric xi_x^2 at the collocated / staggered points of one direction: grid.cc sets xix2 = pow(xi_x, 2.0)
 * element by element, pads included (createUniformGrid, grid.cc:404-421, runs first and fills the whole arrays; the tanh
 * generator, grid.cc:458-476, then overwrites the core).  m1 = the first-derivative metric (oracle/proj_oracle.c).
 * Pinned bit for bit against the reference's grid object in tests/test_impl_oracle.py. */
void implo_metric2(int count, const double *m1, double *out) {
    for (int q = 0; q < count; q++) out[q] = pow(m1[q], 2.0);
}

/* ---- several ranks (npX x npY pencils, all emulated in one process; rank = yRank * npX + xRank) ---------------------------------
 * What changes against one rank: the bulk is cut only on the ranks that own a wall (field.cc:206-212), syncData exchanges faces
 * with the neighbouring ranks (periodic wrap; MPI_PROC_NULL at the walls of a non-periodic direction, hydro.cc:174-203) and
 * edges with the diagonal ranks (always the periodic wrap, parallel.cc:124-129), the wall objects act only where rankFlag is set
 * (boundary.cc:57-73), and maxError is the MPI_MAX over ranks (plainvf.h:97-105).  Pinned like the one-rank version
 * (tests/test_oracle_vs_timestepper.py, 2x1 / 2x2 / 1x2 ranks). */
static void implo_geometry_rank(const implo_params *p, int comp, int npX, int npY, int r, implo_geom *g) {
    const int xr = r % npX, yr = r / npX;
    const int first[3] = {xr == 0, yr == 0, 1}, last[3] = {xr == npX - 1, yr == npY - 1, 1};
    for (int d = 0; d < 3; d++) {
        g->stag[d] = (d != comp);
        g->cs[d] = g->stag[d] ? p->n[d] : p->n[d] - 1;
        g->ext[d] = g->cs[d] + 2;
        const int cut = g->stag[d] && !p->per[d];
        g->lb[d] = (cut && first[d]) ? 1 : 0;
        g->ub[d] = g->cs[d] - 1 - ((cut && last[d]) ? 1 : 0);
    }
}
static int implo_rank_of(int xr, int yr, int npX, int npY) { return ((yr + npY) % npY) * npX + ((xr + npX) % npX); }

/* imposeUBCs / VBCs / WBCs on all ranks: F[r] is rank r's array (all ranks share the extents: equal pencils) */
void implo_impose_bcs_multi(const implo_params *p, int comp, int npX, int npY, double *const *F) {
    const int nr = npX * npY;
    implo_geom G0, *g = &G0;
    implo_geometry_rank(p, comp, npX, npY, 0, g);          /* extents and staggering do not depend on the rank */
    const int ox = g->stag[0] ? 1 : 0, oy = g->stag[1] ? 1 : 0, cx = g->cs[0], cy = g->cs[1], cz = g->cs[2];
    /* faces: every message is read from core points and lands in a pad, so the order does not matter */
    for (int r = 0; r < nr; r++) {
        const int xr = r % npX, yr = r / npX;
        double *D = F[r];
        if (p->per[0] || xr > 0) { const double *S = F[implo_rank_of(xr - 1, yr, npX, npY)];
            for (int j = 0; j < cy; j++) for (int k = 0; k < cz; k++) D[AT(g, -1, j, k)] = S[AT(g, cx - 1 - ox, j, k)]; }
        if (p->per[0] || xr < npX - 1) { const double *S = F[implo_rank_of(xr + 1, yr, npX, npY)];
            for (int j = 0; j < cy; j++) for (int k = 0; k < cz; k++) D[AT(g, cx, j, k)] = S[AT(g, ox, j, k)]; }
        if (p->per[1] || yr > 0) { const double *S = F[implo_rank_of(xr, yr - 1, npX, npY)];
            for (int i = 0; i < cx; i++) for (int k = 0; k < cz; k++) D[AT(g, i, -1, k)] = S[AT(g, i, cy - 1 - oy, k)]; }
        if (p->per[1] || yr < npY - 1) { const double *S = F[implo_rank_of(xr, yr + 1, npX, npY)];
            for (int i = 0; i < cx; i++) for (int k = 0; k < cz; k++) D[AT(g, i, cy, k)] = S[AT(g, i, oy, k)]; }
    }
    for (int r = 0; r < nr; r++) {
        const int xr = r % npX, yr = r / npX;
        double *D = F[r];
        for (int xh = 0; xh < 2; xh++) for (int yh = 0; yh < 2; yh++) {
            const double *S = F[implo_rank_of(xr + (xh ? 1 : -1), yr + (yh ? 1 : -1), npX, npY)];
            for (int k = 0; k < cz; k++) D[AT(g, xh ? cx : -1, yh ? cy : -1, k)] = S[AT(g, xh ? ox : cx - 1 - ox, yh ? oy : cy - 1 - oy, k)];
        }
    }
    /* wall objects: left, right, front, back, top, bottom on the ranks whose rankFlag is set */
    const int order[6] = {0, 1, 2, 3, 5, 4};
    for (int r = 0; r < nr; r++) {
        const int xr = r % npX, yr = r / npX;
        implo_geom GR, *gr = &GR;
        implo_geometry_rank(p, comp, npX, npY, r, gr);
        for (int q = 0; q < 6; q++) {
            const int w = order[q], ax = w / 2, hi = w & 1;
            const int t = p->bc_type[w];
            if (t == IMPL_BC_NONE) continue;
            const int flag = (w == 0) ? (xr == 0) : (w == 1) ? (xr == npX - 1) : (w == 2) ? (yr == 0) : (w == 3) ? (yr == npY - 1) : 1;
            if (!flag && t != IMPL_BC_PERIODIC_CC && t != IMPL_BC_PERIODIC_FC) continue;     /* the periodic classes ignore rankFlag */
            const int wall_lo = gr->lb[ax] - 1, wall_hi = gr->ub[ax] + 1;
            const int dst = hi ? wall_hi : wall_lo;
            int src = dst + (hi ? -1 : 1);
            if (t == IMPL_BC_PERIODIC_CC) src = hi ? wall_lo + 2 : wall_hi - 2;
            if (t == IMPL_BC_PERIODIC_FC) src = hi ? wall_lo + 1 : wall_hi - 1;
            implo_plane(F[r], gr, ax, dst, src, t, p->bc_value[w]);
        }
    }
}

/* solveVx / Vy / Vz on all ranks.  Returns the number of sweeps (the same on every rank: the test is global), -1 on overflow. */
int implo_solve_multi(const implo_params *p, int comp, int npX, int npY, double *const *V, const double *const *rhs,
                      const double *const *m2c_x, const double *const *m2s_x, const double *const *m2c_y, const double *const *m2s_y,
                      double *max_err) {
    const int nr = npX * npY;
    implo_geom G0;
    implo_geometry_rank(p, comp, npX, npY, 0, &G0);
    const size_t tot = (size_t)G0.ext[0] * G0.ext[1] * G0.ext[2];
    double **guess = (double **)malloc(sizeof(double *) * nr);
    for (int r = 0; r < nr; r++) guess[r] = (double *)calloc(tot, sizeof(double));
    const double hx = p->dXi, hy = p->dEt, hz = p->dZt;
    const double hz2hx2 = pow(p->dZt, 2.0) * pow(p->dXi, 2.0), hx2hy2 = pow(p->dXi, 2.0) * pow(p->dEt, 2.0), hy2hz2 = pow(p->dEt, 2.0) * pow(p->dZt, 2.0);
    const double hx2hy2hz2 = pow(p->dXi, 2.0) * pow(p->dEt, 2.0) * pow(p->dZt, 2.0);
    const double inverseRe = 1.0 / p->Re, dt = p->dt, Re = p->Re;
    const long long maxIterations = (long long)(p->n[0] - 1) * (p->n[1] - 1) * (p->n[2] - 1);
    const double *mz = (G0.stag[2] ? p->m2s[2] : p->m2c[2]) + 1;
    int iterCount = 0;
    double maxError = 0.0;
    for (;;) {
        for (int r = 0; r < nr; r++) {
            implo_geom G, *g = &G;
            implo_geometry_rank(p, comp, npX, npY, r, g);
            const double *mx = (g->stag[0] ? m2s_x[r] : m2c_x[r]) + 1, *my = (g->stag[1] ? m2s_y[r] : m2c_y[r]) + 1;
            const double *v = V[r], *b = rhs[r];
            for (int i = g->lb[0]; i <= g->ub[0]; i++) for (int j = g->lb[1]; j <= g->ub[1]; j++) for (int k = g->lb[2]; k <= g->ub[2]; k++)
                guess[r][AT(g, i, j, k)] = ((hy2hz2 * mx[i] * (v[AT(g, i + 1, j, k)] + v[AT(g, i - 1, j, k)]) +
                                             hz2hx2 * my[j] * (v[AT(g, i, j + 1, k)] + v[AT(g, i, j - 1, k)]) +
                                             hx2hy2 * mz[k] * (v[AT(g, i, j, k + 1)] + v[AT(g, i, j, k - 1)])) *
                                            dt / (hx2hy2hz2 * 2.0 * Re) + b[AT(g, i, j, k)]) /
                                           (1.0 + dt * ((hy2hz2 * mx[i] + hz2hx2 * my[j] + hx2hy2 * mz[k])) / (Re * hx2hy2hz2));
        }
        for (int r = 0; r < nr; r++) memcpy(V[r], guess[r], tot * sizeof(double));
        implo_impose_bcs_multi(p, comp, npX, npY, V);
        maxError = 0.0;
        for (int r = 0; r < nr; r++) {
            implo_geom G, *g = &G;
            implo_geometry_rank(p, comp, npX, npY, r, g);
            const double *mx = (g->stag[0] ? m2s_x[r] : m2c_x[r]) + 1, *my = (g->stag[1] ? m2s_y[r] : m2c_y[r]) + 1;
            const double *v = V[r], *b = rhs[r];
            for (int i = g->lb[0]; i <= g->ub[0]; i++) for (int j = g->lb[1]; j <= g->ub[1]; j++) for (int k = g->lb[2]; k <= g->ub[2]; k++) {
                const double c = v[AT(g, i, j, k)];
                const double lap = c - (mx[i] * (v[AT(g, i + 1, j, k)] - 2.0 * c + v[AT(g, i - 1, j, k)]) / (hx * hx) +
                                        my[j] * (v[AT(g, i, j + 1, k)] - 2.0 * c + v[AT(g, i, j - 1, k)]) / (hy * hy) +
                                        mz[k] * (v[AT(g, i, j, k + 1)] - 2.0 * c + v[AT(g, i, j, k - 1)]) / (hz * hz)) *
                                           0.5 * dt * inverseRe;
                const double e = fabs(lap - b[AT(g, i, j, k)]);
                if (e > maxError) maxError = e;
            }
        }
        if (maxError < p->tol) break;
        iterCount += 1;
        if (iterCount > maxIterations) { iterCount = -2; break; }
    }
    for (int r = 0; r < nr; r++) free(guess[r]);
    free(guess);
    if (max_err) *max_err = maxError;
    return iterCount + 1;
}

/* ---- scalar_d3::solveT (src/solvers/scalar/scalar_d3.cc:541-594), one rank ---------------------------------------------------------
 * NOT the velocity expression with other names: the diffusivity kappa multiplies inside (`* dt * kappa / (hx2hy2hz2 * 2.0)`,
 * `1.0 + dt * kappa * (...) / hx2hy2hz2`, `* 0.5 * dt * kappa`), the z metric is ztz2COLLOC although T is staggered in z
 * (scalar_d3.cc:551, 555, 570: replicated, not fixed), the boundary objects are those of scalar::initTBC (scalar.cc:186-229)
 * imposed in the order left, right, front, back, top, bottom (scalar.cc:240-255), and the max runs over the collocated index
 * ranges of plainsf::fxMax (plainsf.h:123-131), i.e. without the last staggered point of a periodic direction.
 * p: comp is ignored (T is staggered in every direction); bc_type / bc_value describe T's walls; Re holds 1 / kappa is NOT
 * used -- kappa is passed explicitly. */
int implo_solve_T(const implo_params *p, double kappa, double *T, const double *rhs, double *max_err) {
    implo_geom G, *g = &G;
    implo_geometry(p, 3, g);                                            /* comp = 3: no collocated direction */
    const size_t tot = (size_t)g->ext[0] * g->ext[1] * g->ext[2];
    double *guess = (double *)calloc(tot, sizeof(double));
    const double hx = p->dXi, hy = p->dEt, hz = p->dZt;
    const double hz2hx2 = pow(p->dZt, 2.0) * pow(p->dXi, 2.0), hx2hy2 = pow(p->dXi, 2.0) * pow(p->dEt, 2.0), hy2hz2 = pow(p->dEt, 2.0) * pow(p->dZt, 2.0);
    const double hx2hy2hz2 = pow(p->dXi, 2.0) * pow(p->dEt, 2.0) * pow(p->dZt, 2.0);
    const double dt = p->dt;
    const double *mx = p->m2s[0] + 1, *my = p->m2s[1] + 1, *mz = p->m2c[2] + 1;       /* ztz2Colloc: indices -1 .. n-1 (n+1 entries) */
    const long long maxIterations = (long long)(p->n[0] - 1) * (p->n[1] - 1) * (p->n[2] - 1);
    int iterCount = 0;
    double maxError = 0.0;
    for (;;) {
        for (int i = g->lb[0]; i <= g->ub[0]; i++) for (int j = g->lb[1]; j <= g->ub[1]; j++) for (int k = g->lb[2]; k <= g->ub[2]; k++)
            guess[AT(g, i, j, k)] = ((hy2hz2 * mx[i] * (T[AT(g, i + 1, j, k)] + T[AT(g, i - 1, j, k)]) +
                                      hz2hx2 * my[j] * (T[AT(g, i, j + 1, k)] + T[AT(g, i, j - 1, k)]) +
                                      hx2hy2 * mz[k] * (T[AT(g, i, j, k + 1)] + T[AT(g, i, j, k - 1)])) *
                                     dt * kappa / (hx2hy2hz2 * 2.0) + rhs[AT(g, i, j, k)]) /
                                    (1.0 + dt * kappa * ((hy2hz2 * mx[i] + hz2hx2 * my[j] + hx2hy2 * mz[k])) / hx2hy2hz2);
        memcpy(T, guess, tot * sizeof(double));                         /* T = guessedScalar (sfield.cc:284-286) */
        {   /* imposeTBCs: the velocity routine with comp = 3 (all staggered) does exactly T.syncData() + the wall objects */
            implo_params q = *p;
            implo_impose_bcs(&q, 3, T);
        }
        maxError = 0.0;
        for (int i = g->lb[0]; i <= g->ub[0]; i++) for (int j = g->lb[1]; j <= g->ub[1]; j++) for (int k = g->lb[2]; k <= g->ub[2]; k++) {
            if (i > p->n[0] - 2 || j > p->n[1] - 2 || k > p->n[2] - 2) continue;        /* fxMax: F(xColl, yColl, zColl) */
            const double c = T[AT(g, i, j, k)];
            const double lap = c - (mx[i] * (T[AT(g, i + 1, j, k)] - 2.0 * c + T[AT(g, i - 1, j, k)]) / (hx * hx) +
                                    my[j] * (T[AT(g, i, j + 1, k)] - 2.0 * c + T[AT(g, i, j - 1, k)]) / (hy * hy) +
                                    mz[k] * (T[AT(g, i, j, k + 1)] - 2.0 * c + T[AT(g, i, j, k - 1)]) / (hz * hz)) *
                                       0.5 * dt * kappa;
            const double e = fabs(lap - rhs[AT(g, i, j, k)]);
            if (e > maxError) maxError = e;
        }
        if (maxError < p->tol) break;
        iterCount += 1;
        if (iterCount > maxIterations) { free(guess); if (max_err) *max_err = maxError; return -1; }
    }
    free(guess);
    if (max_err) *max_err = maxError;
    return iterCount + 1;
}
