/*
 * mg_oracle.c -- CPU parity oracle (plain C restatement) of SARAS's multigrid
 * pressure-Poisson solver.  TEST INFRASTRUCTURE ONLY -- see mg_oracle.h.
 *
 * Compile with strict IEEE semantics:  gcc -O2 -ffp-contract=off  (no -ffast-math).
 * Parity pin: checked bit-for-bit against oracle/_ref (the reference's own sources
 * compiled against stand-in headers) by tests/test_oracle_vs_ref.py.
 *
 * Citations "d3:N" = /root/reference/lib/poisson/poisson_d3.cc line N,
 *           "d2:N" = .../poisson_d2.cc, "p:N" = .../poisson.cc,
 *           "grid:N" = /root/reference/lib/grid/grid.cc, "par:N" = lib/parallel/parallel.cc.
 */
#include "mg_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PROC_NULL (-1)

typedef struct {
    int n[3];       /* core points: xEnd+1, yEnd+1, zEnd+1   (p:336-357) */
    int F[3];       /* padded extents of stagFull            (p:346-349) */
    size_t size;
} level_dims;

typedef struct {
    int rank, xRank, yRank;
    int faceRanks[4];
    double **arr[4];            /* [which][level] -> padded array */
    double **metric[6];         /* [which][level] -> padded 1-D array, index base -1 */
    double *stag[3];            /* x/y/zStaggr, length F_d(0), base -1 */
    double *wall[3];            /* xWall (Fy*Fz), yWall (Fx*Fz), zWall (Fx*Fy); 2-D: xWall (Fz), zWall (Fx) */
} rank_state;

struct mgo_solver {
    mgo_params prm;
    int nProc;
    int vLevel, zeroBC, maxCount;
    level_dims *lev;            /* vcDepth+1 */
    double *coef[MGO_NCOEF];    /* per level */
    double dXi, dEt, dZt, xLen, yLen, zLen;
    rank_state *rk;
    double lups;
    double lastGlobalAvg;
    int exactMean;              /* test mode: local sums of the RHS mean are correctly rounded (exact_sum) instead of sequential */
};

/* ------------------------------------------------------------------------------------------ */
/* index helpers: Blitz row-major, z fastest, base index -1 in every direction                 */
/* ------------------------------------------------------------------------------------------ */
#define IDX(L, i, j, k) ((((size_t)((i) + 1)) * (size_t)(L)->F[1] + (size_t)((j) + 1)) * (size_t)(L)->F[2] + (size_t)((k) + 1))

static int ipow2(int e) { return 1 << e; }

static int pmod(int a, int b) { return (a % b + b) % b; }                       /* parallel.h:104 */
static int find_rank(const mgo_params *p, int xR, int yR) {                     /* parallel.h:121 */
    return pmod(yR, p->npY) * p->npX + pmod(xR, p->npX);
}

static int ilog2(int v) { int r = 0; while ((1 << (r + 1)) <= v) r++; return r; }

/* ------------------------------------------------------------------------------------------ */
/* grid metrics: grid.cc:59-136, 176-215, 333-352, 370-422, 442-519                            */
/* Only the *Staggr arrays the Poisson solver samples (p:467-494) plus the coordinates.         */
/* ------------------------------------------------------------------------------------------ */
static void make_grid_dir(const mgo_solver *s, int dir, int coreStart, int nCore, int gridType, double beta,
                          double len, double dxi, double *coord, double *m1, double *m2, double *m0 /* xi_x */) {
    /* arrays have length nCore+2, base -1: element a[i+1] is index i */
    int F = nCore + 2;
    (void)s; (void)dir;
    /* grid:339-351 xiGlo(i) = real(i)*dXi ; grid:107-109 local slice */
    double *xi = (double *)malloc(sizeof(double) * (size_t)F);
    for (int i = -1; i <= nCore; i++) xi[i + 1] = (double)(coreStart + i) * dxi;

    /* grid:308-320 defaults */
    for (int i = 0; i < F; i++) { coord[i] = 1.0; m0[i] = 1.0; m1[i] = 0.0; m2[i] = 1.0; }

    /* grid:370-422 createUniformGrid: whole arrays (pads included) */
    for (int i = -1; i < nCore + 1; i++) coord[i + 1] = len * xi[i + 1];
    {
        double xi_x = 1.0 / len;
        double xix2 = pow(xi_x, 2.0);
        for (int i = 0; i < F; i++) { m0[i] = xi_x; m2[i] = xix2; }
    }

    if (gridType == 2) {
        /* grid:458-466 (and :479-487, :500-508): core points only, pads keep the uniform values */
        double b = beta;
        for (int i = 0; i < nCore; i++) {
            double x = len * (1.0 - tanh(b * (1.0 - 2.0 * xi[i + 1])) / tanh(b)) / 2.0;
            coord[i + 1] = x;
            m0[i + 1] = tanh(b) / (b * len * (1.0 - pow((1.0 - 2.0 * x / len) * tanh(b), 2)));
            m1[i + 1] = -4.0 * pow(tanh(b), 3) * (1.0 - 2.0 * x / len) /
                        (b * len * len * pow(1.0 - pow(tanh(b) * (1.0 - 2.0 * x / len), 2), 2));
            m2[i + 1] = pow(m0[i + 1], 2.0);
        }
    }
    free(xi);
}

/* ------------------------------------------------------------------------------------------ */
/* construction: p:62-83, 326-362, 377-387, 399-439, 453-496, 293-316; d3:462-572; par:89-131   */
/* ------------------------------------------------------------------------------------------ */
mgo_solver *mgo_create(const mgo_params *p) {
    mgo_solver *s = (mgo_solver *)calloc(1, sizeof(mgo_solver));
    s->prm = *p;
    const int D = p->vcDepth;
    const int planar = (p->dim == 2);
    s->nProc = p->npX * p->npY;

    /* grid:75-92 */
    int gsz[3];
    int ind[3] = {p->xInd, planar ? 0 : p->yInd, p->zInd};
    for (int d = 0; d < 3; d++) gsz[d] = (ind[d] == 0) ? 1 : (ipow2(ind[d]) + 1);   /* grid:155-164 */
    s->xLen = p->Lx; s->yLen = p->Ly; s->zLen = p->Lz;
    s->dXi = 1.0 / (double)(gsz[0] - 1);
    s->dEt = planar ? 1.0 : 1.0 / (double)(gsz[1] - 1);
    s->dZt = 1.0 / (double)(gsz[2] - 1);
    if (planar) { s->yLen = 1.0; }

    /* p:377-387 setLocalSizeIndex */
    int lsi[3];
    lsi[0] = ind[0] - ilog2(p->npX);
    lsi[1] = planar ? ind[1] : ind[1] - ilog2(p->npY);
    lsi[2] = ind[2];

    /* p:326-362 setStagBounds */
    s->lev = (level_dims *)calloc((size_t)D + 1, sizeof(level_dims));
    for (int l = 0; l <= D; l++) {
        level_dims *L = &s->lev[l];
        L->n[0] = ipow2(lsi[0] - l) + 1;
        L->n[1] = planar ? 1 : ipow2(lsi[1] - l) + 1;
        L->n[2] = ipow2(lsi[2] - l) + 1;
        for (int d = 0; d < 3; d++) L->F[d] = L->n[d] + 2;
        L->size = (size_t)L->F[0] * L->F[1] * L->F[2];
    }
    {   /* p:360-361: cgSize = ubound - lbound of stagFull(D) = F-1 */
        level_dims *L = &s->lev[D];
        s->maxCount = (L->F[0] - 1) * (L->F[1] - 1) * (L->F[2] - 1);
    }

    /* p:399-439 setCoefficients */
    for (int c = 0; c < MGO_NCOEF; c++) s->coef[c] = (double *)calloc((size_t)D + 1, sizeof(double));
    for (int l = 0; l <= D; l++) {
        int stride = ipow2(l);                                          /* p:74-77 strideValues */
        s->coef[MGO_HX][l] = stride * s->dXi;
        s->coef[MGO_HY][l] = stride * s->dEt;
        s->coef[MGO_HZ][l] = stride * s->dZt;
        s->coef[MGO_HX2][l] = pow(stride * s->dXi, 2.0);
        s->coef[MGO_HZ2][l] = pow(stride * s->dZt, 2.0);
        s->coef[MGO_HXHY][l] = pow(stride, 4.0) * pow(s->dXi, 2.0) * pow(s->dEt, 2.0);
        s->coef[MGO_HYHZ][l] = pow(stride, 4.0) * pow(s->dEt, 2.0) * pow(s->dZt, 2.0);
        s->coef[MGO_HZHX][l] = pow(stride, 4.0) * pow(s->dZt, 2.0) * pow(s->dXi, 2.0);
        s->coef[MGO_HXHYHZ][l] = pow(stride, 6.0) * pow(s->dXi, 2.0) * pow(s->dEt, 2.0) * pow(s->dZt, 2.0);
    }

    s->rk = (rank_state *)calloc((size_t)s->nProc, sizeof(rank_state));
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        R->rank = r;
        R->xRank = r % p->npX;                                          /* par:89-92 */
        R->yRank = r / p->npX;
        /* par:101-131 getNeighbours (periodic by construction) */
        for (int f = 0; f < 4; f++) R->faceRanks[f] = PROC_NULL;
        R->faceRanks[0] = find_rank(p, R->xRank - 1, R->yRank);
        R->faceRanks[1] = find_rank(p, R->xRank + 1, R->yRank);
        if (!planar) {
            R->faceRanks[2] = find_rank(p, R->xRank, R->yRank - 1);
            R->faceRanks[3] = find_rank(p, R->xRank, R->yRank + 1);
        }
        /* hydro.cc:174-203 checkPeriodic (runs after the multigrid constructor, before the first mgSolve) */
        if (p->wall_null) {
            if (!p->xPer) {
                if (R->xRank == 0) R->faceRanks[0] = PROC_NULL;
                if (R->xRank == p->npX - 1) R->faceRanks[1] = PROC_NULL;
            }
            if (planar) {
                R->faceRanks[2] = PROC_NULL; R->faceRanks[3] = PROC_NULL;
            } else if (!p->yPer) {
                if (R->yRank == 0) R->faceRanks[2] = PROC_NULL;
                if (R->yRank == p->npY - 1) R->faceRanks[3] = PROC_NULL;
            }
        }

        /* p:293-316 initializeArrays */
        for (int a = 0; a < 4; a++) {
            R->arr[a] = (double **)calloc((size_t)D + 1, sizeof(double *));
            for (int l = 0; l <= D; l++) R->arr[a][l] = (double *)calloc(s->lev[l].size, sizeof(double));
        }

        /* grid metrics of this rank on level 0 (grid:176-215 sub-domain start = rank*(localN-1)) */
        level_dims *L0 = &s->lev[0];
        double *g1[3], *g2[3], *g0[3];
        int gtype[3] = {p->xGrid, planar ? 0 : p->yGrid, p->zGrid};
        double beta[3] = {p->betaX, p->betaY, p->betaZ};
        double len[3] = {s->xLen, s->yLen, s->zLen};
        double dxi[3] = {s->dXi, s->dEt, s->dZt};
        int start[3] = {R->xRank * (L0->n[0] - 1), planar ? 0 : R->yRank * (L0->n[1] - 1), 0};
        for (int d = 0; d < 3; d++) {
            int F = L0->F[d];
            R->stag[d] = (double *)calloc((size_t)F, sizeof(double));
            g0[d] = (double *)calloc((size_t)F, sizeof(double));
            g1[d] = (double *)calloc((size_t)F, sizeof(double));
            g2[d] = (double *)calloc((size_t)F, sizeof(double));
            make_grid_dir(s, d, start[d], L0->n[d], gtype[d], beta[d], len[d], dxi[d], R->stag[d], g1[d], g2[d], g0[d]);
            if (planar && d == 1) {
                /* PLANAR: the y arrays are never generated (grid:383-393 skipped), only defaults and the
                 * uniform metric assignment (grid:411-415) apply; unused by multigrid_d2 */
            }
        }
        /* p:453-496 copyStaggrDerivs: level arrays zero, core sampled with stride 2^l */
        for (int m = 0; m < 6; m++) {
            int d = m / 2;
            R->metric[m] = (double **)calloc((size_t)D + 1, sizeof(double *));
            for (int l = 0; l <= D; l++) {
                level_dims *L = &s->lev[l];
                double *dst = (double *)calloc((size_t)L->F[d], sizeof(double));
                const double *src = (m % 2 == 0) ? g1[d] : g2[d];
                int stride = ipow2(l);
                if (!(planar && d == 1)) {
                    for (int i = 0; i < L->n[d]; i++) dst[i + 1] = src[i * stride + 1];
                }
                R->metric[m][l] = dst;
            }
        }
        for (int d = 0; d < 3; d++) { free(g0[d]); free(g1[d]); free(g2[d]); }

        /* d3:513-572 / d2:367-400 initDirichlet (allocated in every mode; only used under TEST_POISSON) */
        {
            double hx0 = s->coef[MGO_HX][0], hy0 = s->coef[MGO_HY][0], hz0 = s->coef[MGO_HZ][0];
            int ubx = L0->n[0] - 1, uby = L0->n[1] - 1, ubz = L0->n[2] - 1;
            int halfIndX = ubx * p->npX / 2;
            int halfIndY = uby * p->npY / 2;
            double xDist, yDist, zDist;
            if (!planar) {
                R->wall[0] = (double *)calloc((size_t)L0->F[1] * L0->F[2], sizeof(double));
                R->wall[1] = (double *)calloc((size_t)L0->F[0] * L0->F[2], sizeof(double));
                R->wall[2] = (double *)calloc((size_t)L0->F[0] * L0->F[1], sizeof(double));
                xDist = hx0 + p->Lx / 2.0;
                for (int j = 0; j <= uby; ++j) {
                    yDist = hy0 * (R->yRank * uby + j - halfIndY);
                    for (int k = 0; k <= ubz; ++k) {
                        zDist = hz0 * (k - ubz / 2);
                        R->wall[0][(size_t)(j + 1) * L0->F[2] + (k + 1)] = (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0;
                    }
                }
                yDist = hy0 + p->Ly / 2.0;
                for (int i = 0; i <= ubx; ++i) {
                    xDist = hx0 * (R->xRank * ubx + i - halfIndX);
                    for (int k = 0; k <= ubz; ++k) {
                        zDist = hz0 * (k - ubz / 2);
                        R->wall[1][(size_t)(i + 1) * L0->F[2] + (k + 1)] = (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0;
                    }
                }
                zDist = hz0 + p->Lz / 2.0;
                for (int i = 0; i <= ubx; ++i) {
                    xDist = hx0 * (R->xRank * ubx + i - halfIndX);
                    for (int j = 0; j <= uby; ++j) {
                        yDist = hy0 * (R->yRank * uby + j - halfIndY);
                        R->wall[2][(size_t)(i + 1) * L0->F[1] + (j + 1)] = (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0;
                    }
                }
            } else {
                R->wall[0] = (double *)calloc((size_t)L0->F[2], sizeof(double));
                R->wall[1] = NULL;
                R->wall[2] = (double *)calloc((size_t)L0->F[0], sizeof(double));
                xDist = hx0 + p->Lx / 2.0;
                for (int k = 0; k <= ubz; ++k) {
                    zDist = hz0 * (k - ubz / 2);
                    R->wall[0][k + 1] = (xDist * xDist + zDist * zDist) / 4.0;
                }
                zDist = hz0 + p->Lz / 2.0;
                for (int i = 0; i <= ubx; ++i) {
                    xDist = hx0 * (R->xRank * ubx + i - halfIndX);
                    R->wall[2][i + 1] = (xDist * xDist + zDist * zDist) / 4.0;
                }
            }
        }
    }
    return s;
}

void mgo_destroy(mgo_solver *s) {
    if (!s) return;
    const int D = s->prm.vcDepth;
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        for (int a = 0; a < 4; a++) { for (int l = 0; l <= D; l++) free(R->arr[a][l]); free(R->arr[a]); }
        for (int m = 0; m < 6; m++) { for (int l = 0; l <= D; l++) free(R->metric[m][l]); free(R->metric[m]); }
        for (int d = 0; d < 3; d++) { free(R->stag[d]); free(R->wall[d]); }
    }
    free(s->rk);
    for (int c = 0; c < MGO_NCOEF; c++) free(s->coef[c]);
    free(s->lev);
    free(s);
}

int mgo_nranks(const mgo_solver *s) { return s->nProc; }
void mgo_level_dims(const mgo_solver *s, int level, int n[3], int F[3]) {
    for (int d = 0; d < 3; d++) { n[d] = s->lev[level].n[d]; F[d] = s->lev[level].F[d]; }
}
void mgo_rank_info(const mgo_solver *s, int rank, int out[6]) {
    out[0] = s->rk[rank].xRank; out[1] = s->rk[rank].yRank;
    for (int f = 0; f < 4; f++) out[2 + f] = s->rk[rank].faceRanks[f];
}
void mgo_metric(const mgo_solver *s, int rank, int level, int which, double *out) {
    memcpy(out, s->rk[rank].metric[which][level], sizeof(double) * (size_t)s->lev[level].F[which / 2]);
}
double mgo_coef(const mgo_solver *s, int level, int which) { return s->coef[which][level]; }
void mgo_grid_spacing(const mgo_solver *s, double out[3]) { out[0] = s->dXi; out[1] = s->dEt; out[2] = s->dZt; }
void mgo_staggr_coords(const mgo_solver *s, int rank, int dir, double *out) {
    memcpy(out, s->rk[rank].stag[dir], sizeof(double) * (size_t)s->lev[0].F[dir]);
}
void mgo_wall(const mgo_solver *s, int rank, int which, double *out) {
    const level_dims *L0 = &s->lev[0];
    size_t n;
    if (s->prm.dim == 3) n = (which == 0) ? (size_t)L0->F[1] * L0->F[2] : (which == 1) ? (size_t)L0->F[0] * L0->F[2] : (size_t)L0->F[0] * L0->F[1];
    else n = (which == 0) ? (size_t)L0->F[2] : (which == 2) ? (size_t)L0->F[0] : 0;
    if (n && s->rk[rank].wall[which]) memcpy(out, s->rk[rank].wall[which], sizeof(double) * n);
}
int mgo_max_count(const mgo_solver *s) { return s->maxCount; }
double *mgo_array(mgo_solver *s, int rank, int level, int which) { return s->rk[rank].arr[which][level]; }
void mgo_set_level(mgo_solver *s, int v) { s->vLevel = v; }
int mgo_get_level(const mgo_solver *s) { return s->vLevel; }
void mgo_set_zero_bc(mgo_solver *s, int z) { s->zeroBC = z; }
double mgo_lup_count(const mgo_solver *s) { return s->lups; }
double mgo_last_global_avg(const mgo_solver *s) { return s->lastGlobalAvg; }
void mgo_set_exact_mean(mgo_solver *s, int on) { s->exactMean = on; }

/* Correctly rounded sum of n doubles (Shewchuk's non-overlapping partials, the algorithm behind Python's math.fsum).
 * NOT the reference's arithmetic: poisson.cc:122 is one running fp64 sum.  It exists so that a test can show that the
 * device's compensated sum is the exact mean, i.e. that the production-mode field difference between the CUDA path and the
 * reference is the rounding error of the REFERENCE's sequential sum and nothing else (tests/test_gpu_parity.py). */
static double exact_sum(const double *a, size_t n) {
    double part[64];
    int np = 0;
    for (size_t q = 0; q < n; q++) {
        double x = a[q];
        int i = 0;
        for (int j = 0; j < np; j++) {
            double y = part[j];
            if (fabs(x) < fabs(y)) { const double t = x; x = y; y = t; }
            const double hi = x + y;
            const double lo = y - (hi - x);
            if (lo != 0.0) part[i++] = lo;
            x = hi;
        }
        np = i;
        if (np >= 63) {                                     /* cannot happen for finite doubles (<= 40 partials); stay safe */
            double t = x;
            for (int j = 0; j < np; j++) t += part[j];
            x = t; np = 0;
        }
        part[np++] = x;
    }
    double hi = 0.0, lo = 0.0;
    int m = np;
    if (m > 0) {
        hi = part[--m];
        while (m > 0) {
            const double x = hi, y = part[--m];
            hi = x + y;
            const double yr = hi - x;
            lo = y - yr;
            if (lo != 0.0) break;
        }
        /* round-half-even correction when the discarded tail points the same way as lo */
        if (m > 0 && ((lo < 0.0 && part[m - 1] < 0.0) || (lo > 0.0 && part[m - 1] > 0.0))) {
            const double y = lo * 2.0, x = hi + y, yr = x - hi;
            if (y == yr) hi = x;
        }
    }
    return hi;
}
void mgo_lup_reset(mgo_solver *s) { s->lups = 0.0; }

/* ------------------------------------------------------------------------------------------ */
/* updatePads: d3:671-686 (+ datatypes d3:462-510), d2:467-478 (+ d2:341-364)                   */
/*                                                                                              */
/* x faces: one contiguous block of Fy*Fz doubles (3-D) / Fz doubles on row j=0 (2-D) from plane */
/* i=1 to the left neighbour's i=xEnd+1 and from i=xEnd-1 to the right neighbour's i=-1.         */
/* y faces: Fx blocks of Fz doubles, stride Fy*Fz, from row j=1 / j=yEnd-1.                       */
/* The reference posts all four receives, then does four blocking sends in the order             */
/* left, right, front, back.  With one rank (self sends) this executes strictly in program order: */
/* x faces land before the y rows (which include the x pads) are read.  With several ranks the    */
/* edge pads are racy in the reference; the oracle fixes the same order: all x transfers of all   */
/* ranks complete, then all y transfers.  PROC_NULL neighbours transfer nothing.                  */
/* ------------------------------------------------------------------------------------------ */
void mgo_update_pads(mgo_solver *s, int which) {
    const int l = s->vLevel;
    const level_dims *L = &s->lev[l];
    const int planar = (s->prm.dim == 2);
    const int xE = L->n[0] - 1, yE = L->n[1] - 1;
    const size_t plane = (size_t)L->F[1] * L->F[2];

    /* phase X.  Snapshot the send planes first so that simultaneous sends do not see received data
     * (send planes i=1, i=xEnd-1 and receive planes i=-1, i=xEnd+1 never overlap, so this is exact). */
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        double *me = R->arr[which][l];
        /* my left pad i=-1 is written by my left neighbour's "send right" (tag 1): its plane i=xEnd-1 */
        if (R->faceRanks[0] != PROC_NULL) {
            const double *src = s->rk[R->faceRanks[0]].arr[which][l];
            if (!planar) memcpy(&me[IDX(L, -1, -1, -1)], &src[IDX(L, xE - 1, -1, -1)], plane * sizeof(double));
            else memcpy(&me[IDX(L, -1, 0, -1)], &src[IDX(L, xE - 1, 0, -1)], (size_t)L->F[2] * sizeof(double));
        }
        /* my right pad i=xEnd+1 is written by my right neighbour's "send left" (tag 2): its plane i=1 */
        if (R->faceRanks[1] != PROC_NULL) {
            const double *src = s->rk[R->faceRanks[1]].arr[which][l];
            if (!planar) memcpy(&me[IDX(L, xE + 1, -1, -1)], &src[IDX(L, 1, -1, -1)], plane * sizeof(double));
            else memcpy(&me[IDX(L, xE + 1, 0, -1)], &src[IDX(L, 1, 0, -1)], (size_t)L->F[2] * sizeof(double));
        }
    }
    if (planar) return;

    /* phase Y: rows j=1 / j=yEnd-1, all i in [-1, xEnd+1], all k */
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        double *me = R->arr[which][l];
        if (R->faceRanks[2] != PROC_NULL) {
            const double *src = s->rk[R->faceRanks[2]].arr[which][l];
            for (int i = -1; i <= xE + 1; i++)
                memcpy(&me[IDX(L, i, -1, -1)], &src[IDX(L, i, yE - 1, -1)], (size_t)L->F[2] * sizeof(double));
        }
        if (R->faceRanks[3] != PROC_NULL) {
            const double *src = s->rk[R->faceRanks[3]].arr[which][l];
            for (int i = -1; i <= xE + 1; i++)
                memcpy(&me[IDX(L, i, yE + 1, -1)], &src[IDX(L, i, 1, -1)], (size_t)L->F[2] * sizeof(double));
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* imposeBC: d3:575-668, d2:403-464                                                            */
/* ------------------------------------------------------------------------------------------ */
static void impose_bc_rank(mgo_solver *s, rank_state *R) {
    const int l = s->vLevel;
    const level_dims *L = &s->lev[l];
    const mgo_params *p = &s->prm;
    double *P = R->arr[MGO_PRESSURE][l];
    const int xE = L->n[0] - 1, yE = L->n[1] - 1, zE = L->n[2] - 1;
    const int Fx = L->F[0], Fy = L->F[1], Fz = L->F[2];
    const int planar = (p->dim == 2);
    /* in 2-D only row j=0 is touched (d2 uses (.., 0, all)); in 3-D full planes (all, all) */
    const int jlo = planar ? 0 : -1, jhi = planar ? 0 : Fy - 2;

    if (!p->xPer) {
        if (p->test_poisson) {
            if (s->zeroBC) {
                if (R->xRank == 0)
                    for (int j = jlo; j <= jhi; j++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, -1, j, k)] = -P[IDX(L, 1, j, k)];
                if (R->xRank == p->npX - 1)
                    for (int j = jlo; j <= jhi; j++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, xE + 1, j, k)] = -P[IDX(L, xE - 1, j, k)];
            } else {
                /* xWall exists at level-0 extents only; the reference only reaches this branch at vLevel 0 */
                const double *W = R->wall[0];
                if (R->xRank == 0)
                    for (int j = jlo; j <= jhi; j++) for (int k = -1; k <= Fz - 2; k++)
                        P[IDX(L, -1, j, k)] = planar ? W[k + 1] : W[(size_t)(j + 1) * Fz + (k + 1)];
                if (R->xRank == p->npX - 1)
                    for (int j = jlo; j <= jhi; j++) for (int k = -1; k <= Fz - 2; k++)
                        P[IDX(L, xE + 1, j, k)] = planar ? W[k + 1] : W[(size_t)(j + 1) * Fz + (k + 1)];
            }
        } else {
            if (R->xRank == 0)
                for (int j = jlo; j <= jhi; j++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, -1, j, k)] = P[IDX(L, 1, j, k)];
            if (R->xRank == p->npX - 1)
                for (int j = jlo; j <= jhi; j++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, xE + 1, j, k)] = P[IDX(L, xE - 1, j, k)];
        }
    }

    if (!planar && !p->yPer) {
        if (p->test_poisson) {
            if (s->zeroBC) {
                if (R->yRank == 0)
                    for (int i = -1; i <= Fx - 2; i++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, i, -1, k)] = -P[IDX(L, i, 1, k)];
                if (R->yRank == p->npY - 1)
                    for (int i = -1; i <= Fx - 2; i++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, i, yE + 1, k)] = -P[IDX(L, i, yE - 1, k)];
            } else {
                const double *W = R->wall[1];
                if (R->yRank == 0)
                    for (int i = -1; i <= Fx - 2; i++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, i, -1, k)] = W[(size_t)(i + 1) * Fz + (k + 1)];
                if (R->yRank == p->npY - 1)
                    for (int i = -1; i <= Fx - 2; i++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, i, yE + 1, k)] = W[(size_t)(i + 1) * Fz + (k + 1)];
            }
        } else {
            if (R->yRank == 0)
                for (int i = -1; i <= Fx - 2; i++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, i, -1, k)] = P[IDX(L, i, 1, k)];
            if (R->yRank == p->npY - 1)
                for (int i = -1; i <= Fx - 2; i++) for (int k = -1; k <= Fz - 2; k++) P[IDX(L, i, yE + 1, k)] = P[IDX(L, i, yE - 1, k)];
        }
    }

    if (p->zPer) {
        for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++) P[IDX(L, i, j, -1)] = P[IDX(L, i, j, zE - 1)];
        for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++) P[IDX(L, i, j, zE + 1)] = P[IDX(L, i, j, 1)];
    } else if (p->test_poisson) {
        if (s->zeroBC) {
            for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++) P[IDX(L, i, j, -1)] = -P[IDX(L, i, j, 1)];
            for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++) P[IDX(L, i, j, zE + 1)] = -P[IDX(L, i, j, zE - 1)];
        } else {
            const double *W = R->wall[2];
            for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++)
                P[IDX(L, i, j, -1)] = planar ? W[i + 1] : W[(size_t)(i + 1) * Fy + (j + 1)];
            for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++)
                P[IDX(L, i, j, zE + 1)] = planar ? W[i + 1] : W[(size_t)(i + 1) * Fy + (j + 1)];
        }
    } else {
        for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++) P[IDX(L, i, j, -1)] = P[IDX(L, i, j, 1)];
        for (int i = -1; i <= Fx - 2; i++) for (int j = jlo; j <= jhi; j++) P[IDX(L, i, j, zE + 1)] = P[IDX(L, i, j, zE - 1)];
    }
}

void mgo_impose_bc(mgo_solver *s) {
    mgo_update_pads(s, MGO_PRESSURE);                       /* d3:576 */
    for (int r = 0; r < s->nProc; r++) impose_bc_rank(s, &s->rk[r]);
}

/* ------------------------------------------------------------------------------------------ */
/* stencils.  The expressions are transcribed operand for operand so that, compiled with         */
/* -ffp-contract=off, every rounding happens where the reference source puts it.                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const double *xixx, *xix2, *etyy, *ety2, *ztzz, *ztz2;  /* base pointers shifted so [i] is index i */
    double hx, hy, hz, hx2, hz2, hzhx, hxhy, hyhz, hxhyhz;
} stencil_ctx;

static stencil_ctx make_ctx(const mgo_solver *s, const rank_state *R, int l) {
    stencil_ctx c;
    c.xixx = R->metric[MGO_XIXX][l] + 1; c.xix2 = R->metric[MGO_XIX2][l] + 1;
    c.etyy = R->metric[MGO_ETYY][l] + 1; c.ety2 = R->metric[MGO_ETY2][l] + 1;
    c.ztzz = R->metric[MGO_ZTZZ][l] + 1; c.ztz2 = R->metric[MGO_ZTZ2][l] + 1;
    c.hx = s->coef[MGO_HX][l]; c.hy = s->coef[MGO_HY][l]; c.hz = s->coef[MGO_HZ][l];
    c.hx2 = s->coef[MGO_HX2][l]; c.hz2 = s->coef[MGO_HZ2][l]; c.hzhx = s->coef[MGO_HZHX][l];
    c.hxhy = s->coef[MGO_HXHY][l]; c.hyhz = s->coef[MGO_HYHZ][l]; c.hxhyhz = s->coef[MGO_HXHYHZ][l];
    return c;
}

/* d3:140-147 (= :156-163 = :201-208) */
static inline double relax3(const stencil_ctx *c, const level_dims *L, const double *P, const double *Rh, int i, int j, int k) {
    return (c->hyhz * c->xix2[i] * (P[IDX(L, i + 1, j, k)] + P[IDX(L, i - 1, j, k)]) * 2.0 +
            c->hyhz * c->xixx[i] * (P[IDX(L, i + 1, j, k)] - P[IDX(L, i - 1, j, k)]) * c->hx +
            c->hzhx * c->ety2[j] * (P[IDX(L, i, j + 1, k)] + P[IDX(L, i, j - 1, k)]) * 2.0 +
            c->hzhx * c->etyy[j] * (P[IDX(L, i, j + 1, k)] - P[IDX(L, i, j - 1, k)]) * c->hy +
            c->hxhy * c->ztz2[k] * (P[IDX(L, i, j, k + 1)] + P[IDX(L, i, j, k - 1)]) * 2.0 +
            c->hxhy * c->ztzz[k] * (P[IDX(L, i, j, k + 1)] - P[IDX(L, i, j, k - 1)]) * c->hz -
            2.0 * c->hxhyhz * Rh[IDX(L, i, j, k)]) /
           (4.0 * (c->hyhz * c->xix2[i] + c->hzhx * c->ety2[j] + c->hxhy * c->ztz2[k]));
}

/* d2:121-126 (= :133-138 = :160-165) */
static inline double relax2(const stencil_ctx *c, const level_dims *L, const double *P, const double *Rh, int i, int k) {
    return (c->hz2 * c->xix2[i] * (P[IDX(L, i + 1, 0, k)] + P[IDX(L, i - 1, 0, k)]) * 2.0 +
            c->hz2 * c->xixx[i] * (P[IDX(L, i + 1, 0, k)] - P[IDX(L, i - 1, 0, k)]) * c->hx +
            c->hx2 * c->ztz2[k] * (P[IDX(L, i, 0, k + 1)] + P[IDX(L, i, 0, k - 1)]) * 2.0 +
            c->hx2 * c->ztzz[k] * (P[IDX(L, i, 0, k + 1)] - P[IDX(L, i, 0, k - 1)]) * c->hz -
            2.0 * c->hzhx * Rh[IDX(L, i, 0, k)]) /
           (4.0 * (c->hz2 * c->xix2[i] + c->hx2 * c->ztz2[k]));
}

/* the Laplacian bracket of d3:100-105 (= :219-224, :411-416) */
static inline double lap3(const stencil_ctx *c, const level_dims *L, const double *P, int i, int j, int k) {
    return (c->xix2[i] * (P[IDX(L, i + 1, j, k)] - 2.0 * P[IDX(L, i, j, k)] + P[IDX(L, i - 1, j, k)]) / (c->hx * c->hx) +
            c->xixx[i] * (P[IDX(L, i + 1, j, k)] - P[IDX(L, i - 1, j, k)]) / (2.0 * c->hx) +
            c->ety2[j] * (P[IDX(L, i, j + 1, k)] - 2.0 * P[IDX(L, i, j, k)] + P[IDX(L, i, j - 1, k)]) / (c->hy * c->hy) +
            c->etyy[j] * (P[IDX(L, i, j + 1, k)] - P[IDX(L, i, j - 1, k)]) / (2.0 * c->hy) +
            c->ztz2[k] * (P[IDX(L, i, j, k + 1)] - 2.0 * P[IDX(L, i, j, k)] + P[IDX(L, i, j, k - 1)]) / (c->hz * c->hz) +
            c->ztzz[k] * (P[IDX(L, i, j, k + 1)] - P[IDX(L, i, j, k - 1)]) / (2.0 * c->hz));
}

/* d2:98-101 (= :174-177): divides by hx*hx, hz*hz */
static inline double lap2(const stencil_ctx *c, const level_dims *L, const double *P, int i, int k) {
    return (c->xix2[i] * (P[IDX(L, i + 1, 0, k)] - 2.0 * P[IDX(L, i, 0, k)] + P[IDX(L, i - 1, 0, k)]) / (c->hx * c->hx) +
            c->xixx[i] * (P[IDX(L, i + 1, 0, k)] - P[IDX(L, i - 1, 0, k)]) / (2.0 * c->hx) +
            c->ztz2[k] * (P[IDX(L, i, 0, k + 1)] - 2.0 * P[IDX(L, i, 0, k)] + P[IDX(L, i, 0, k - 1)]) / (c->hz * c->hz) +
            c->ztzz[k] * (P[IDX(L, i, 0, k + 1)] - P[IDX(L, i, 0, k - 1)]) / (2.0 * c->hz));
}

/* d2:293-296: computeError in 2-D divides by hx2(0), hz2(0) (pow-computed squares), unlike d2:98-101 */
static inline double lap2_err(const stencil_ctx *c, const level_dims *L, const double *P, int i, int k) {
    return (c->xix2[i] * (P[IDX(L, i + 1, 0, k)] - 2.0 * P[IDX(L, i, 0, k)] + P[IDX(L, i - 1, 0, k)]) / c->hx2 +
            c->xixx[i] * (P[IDX(L, i + 1, 0, k)] - P[IDX(L, i - 1, 0, k)]) / (2.0 * c->hx) +
            c->ztz2[k] * (P[IDX(L, i, 0, k + 1)] - 2.0 * P[IDX(L, i, 0, k)] + P[IDX(L, i, 0, k - 1)]) / c->hz2 +
            c->ztzz[k] * (P[IDX(L, i, 0, k + 1)] - P[IDX(L, i, 0, k - 1)]) / (2.0 * c->hz));
}

/* ------------------------------------------------------------------------------------------ */
/* computeResidual: d3:90-111, d2:90-106                                                       */
/* ------------------------------------------------------------------------------------------ */
void mgo_compute_residual(mgo_solver *s) {
    const int l = s->vLevel;
    const level_dims *L = &s->lev[l];
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        stencil_ctx c = make_ctx(s, R, l);
        double *T = R->arr[MGO_TMP][l];
        const double *P = R->arr[MGO_PRESSURE][l], *Rh = R->arr[MGO_RESIDUAL][l];
        memset(T, 0, L->size * sizeof(double));
        if (s->prm.dim == 3) {
            for (int i = 0; i < L->n[0]; ++i) for (int j = 0; j < L->n[1]; ++j) for (int k = 0; k < L->n[2]; ++k)
                T[IDX(L, i, j, k)] = Rh[IDX(L, i, j, k)] - lap3(&c, L, P, i, j, k);
        } else {
            for (int i = 0; i < L->n[0]; ++i) for (int k = 0; k < L->n[2]; ++k)
                T[IDX(L, i, 0, k)] = Rh[IDX(L, i, 0, k)] - lap2(&c, L, P, i, k);
        }
        s->lups += (double)L->n[0] * L->n[1] * L->n[2];
    }
    mgo_update_pads(s, MGO_TMP);
}

/* ------------------------------------------------------------------------------------------ */
/* smooth: d3:114-187, d2:109-147.  Jacobi: swap(tmpDataArray, pressureData) swaps the outer     */
/* arrays, i.e. every level at once (d3:168) -- reproduced literally.                            */
/* ------------------------------------------------------------------------------------------ */
void mgo_smooth(mgo_solver *s, int count) {
    const int l = s->vLevel;
    const level_dims *L = &s->lev[l];
    for (int r = 0; r < s->nProc; r++) memset(s->rk[r].arr[MGO_TMP][l], 0, L->size * sizeof(double));

    for (int n = 0; n < count; ++n) {
        mgo_impose_bc(s);
        for (int r = 0; r < s->nProc; r++) {
            rank_state *R = &s->rk[r];
            stencil_ctx c = make_ctx(s, R, l);
            double *P = R->arr[MGO_PRESSURE][l];
            double *T = R->arr[MGO_TMP][l];
            const double *Rh = R->arr[MGO_RESIDUAL][l];
            if (s->prm.gsSmooth) {
                if (s->prm.dim == 3) {
                    for (int i = 0; i < L->n[0]; ++i) for (int j = 0; j < L->n[1]; ++j) for (int k = 0; k < L->n[2]; ++k)
                        P[IDX(L, i, j, k)] = relax3(&c, L, P, Rh, i, j, k);
                } else {
                    for (int i = 0; i < L->n[0]; ++i) for (int k = 0; k < L->n[2]; ++k)
                        P[IDX(L, i, 0, k)] = relax2(&c, L, P, Rh, i, k);
                }
            } else {
                if (s->prm.dim == 3) {
                    for (int i = 0; i < L->n[0]; ++i) for (int j = 0; j < L->n[1]; ++j) for (int k = 0; k < L->n[2]; ++k)
                        T[IDX(L, i, j, k)] = relax3(&c, L, P, Rh, i, j, k);
                } else {
                    for (int i = 0; i < L->n[0]; ++i) for (int k = 0; k < L->n[2]; ++k)
                        T[IDX(L, i, 0, k)] = relax2(&c, L, P, Rh, i, k);
                }
                double **tmp = R->arr[MGO_TMP]; R->arr[MGO_TMP] = R->arr[MGO_PRESSURE]; R->arr[MGO_PRESSURE] = tmp;
            }
            s->lups += (double)L->n[0] * L->n[1] * L->n[2];
        }
    }
    mgo_impose_bc(s);
}

/* ------------------------------------------------------------------------------------------ */
/* solve (coarsest level, always Gauss-Seidel): d3:190-247, d2:150-199                          */
/* ------------------------------------------------------------------------------------------ */
int mgo_solve_coarsest(mgo_solver *s) {
    const int l = s->vLevel;
    const level_dims *L = &s->lev[l];
    int iterCount = 0;
    while (1) {
        mgo_impose_bc(s);
        double globalMax = -HUGE_VAL;
        for (int r = 0; r < s->nProc; r++) {
            rank_state *R = &s->rk[r];
            stencil_ctx c = make_ctx(s, R, l);
            double *P = R->arr[MGO_PRESSURE][l];
            const double *Rh = R->arr[MGO_RESIDUAL][l];
            double localMax = -1.0e-10, tempValue;
            if (s->prm.dim == 3) {
                for (int i = 0; i < L->n[0]; ++i) for (int j = 0; j < L->n[1]; ++j) for (int k = 0; k < L->n[2]; ++k)
                    P[IDX(L, i, j, k)] = relax3(&c, L, P, Rh, i, j, k);
                for (int i = 0; i < L->n[0]; ++i) for (int j = 0; j < L->n[1]; ++j) for (int k = 0; k < L->n[2]; ++k) {
                    tempValue = fabs(Rh[IDX(L, i, j, k)] - lap3(&c, L, P, i, j, k));
                    if (tempValue > localMax) localMax = tempValue;
                }
            } else {
                for (int i = 0; i < L->n[0]; ++i) for (int k = 0; k < L->n[2]; ++k)
                    P[IDX(L, i, 0, k)] = relax2(&c, L, P, Rh, i, k);
                for (int i = 0; i < L->n[0]; ++i) for (int k = 0; k < L->n[2]; ++k) {
                    tempValue = fabs(Rh[IDX(L, i, 0, k)] - lap2(&c, L, P, i, k));
                    if (tempValue > localMax) localMax = tempValue;
                }
            }
            s->lups += 2.0 * (double)L->n[0] * L->n[1] * L->n[2];
            if (localMax > globalMax) globalMax = localMax;     /* MPI_Allreduce MAX d3:233 */
        }
        if (globalMax < s->prm.mgTolerance) break;
        iterCount += 1;
        if (iterCount > s->maxCount) break;                     /* warning only when printResidual, d3:240-243 */
    }
    mgo_impose_bc(s);
    return iterCount;
}

/* ------------------------------------------------------------------------------------------ */
/* coarsen: d3:250-298, d2:202-236                                                             */
/* ------------------------------------------------------------------------------------------ */
void mgo_coarsen(mgo_solver *s) {
    const int pLevel = s->vLevel;
    s->vLevel += 1;
    const level_dims *Lf = &s->lev[pLevel];
    const level_dims *Lc = &s->lev[s->vLevel];
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        const double *T = R->arr[MGO_TMP][pLevel];
        double *Rc = R->arr[MGO_RESIDUAL][s->vLevel];
        if (s->prm.dim == 3) {
            for (int i = 0; i < Lc->n[0]; ++i) {
                int i2 = i * 2;
                for (int j = 0; j < Lc->n[1]; ++j) {
                    int j2 = j * 2;
                    for (int k = 0; k < Lc->n[2]; ++k) {
                        int k2 = k * 2;
                        double facePoints = (T[IDX(Lf, i2 + 1, j2, k2)] + T[IDX(Lf, i2 - 1, j2, k2)] +
                                             T[IDX(Lf, i2, j2 + 1, k2)] + T[IDX(Lf, i2, j2 - 1, k2)] +
                                             T[IDX(Lf, i2, j2, k2 + 1)] + T[IDX(Lf, i2, j2, k2 - 1)]) * 0.0625;
                        double edgePoints = (T[IDX(Lf, i2 + 1, j2 + 1, k2)] + T[IDX(Lf, i2 + 1, j2 - 1, k2)] +
                                             T[IDX(Lf, i2 - 1, j2 - 1, k2)] + T[IDX(Lf, i2 - 1, j2 + 1, k2)] +
                                             T[IDX(Lf, i2, j2 + 1, k2 + 1)] + T[IDX(Lf, i2, j2 - 1, k2 + 1)] +
                                             T[IDX(Lf, i2, j2 - 1, k2 - 1)] + T[IDX(Lf, i2, j2 + 1, k2 - 1)] +
                                             T[IDX(Lf, i2 + 1, j2, k2 + 1)] + T[IDX(Lf, i2 + 1, j2, k2 - 1)] +
                                             T[IDX(Lf, i2 - 1, j2, k2 - 1)] + T[IDX(Lf, i2 - 1, j2, k2 + 1)]) * 0.03125;
                        double vertPoints = (T[IDX(Lf, i2 + 1, j2 + 1, k2 + 1)] +
                                             T[IDX(Lf, i2 + 1, j2 + 1, k2 - 1)] +
                                             T[IDX(Lf, i2 + 1, j2 - 1, k2 + 1)] +
                                             T[IDX(Lf, i2 - 1, j2 + 1, k2 + 1)] +
                                             T[IDX(Lf, i2 + 1, j2 - 1, k2 - 1)] +
                                             T[IDX(Lf, i2 - 1, j2 + 1, k2 - 1)] +
                                             T[IDX(Lf, i2 - 1, j2 - 1, k2 + 1)] +
                                             T[IDX(Lf, i2 - 1, j2 - 1, k2 - 1)]) * 0.015625;
                        Rc[IDX(Lc, i, j, k)] = facePoints + edgePoints + vertPoints + T[IDX(Lf, i2, j2, k2)] * 0.125;
                    }
                }
            }
        } else {
            for (int i = 0; i < Lc->n[0]; ++i) {
                int i2 = i * 2;
                for (int k = 0; k < Lc->n[2]; ++k) {
                    int k2 = k * 2;
                    double facePoints = (T[IDX(Lf, i2 + 1, 0, k2)] + T[IDX(Lf, i2 - 1, 0, k2)] +
                                         T[IDX(Lf, i2, 0, k2 + 1)] + T[IDX(Lf, i2, 0, k2 - 1)]) * 0.125;
                    double vertPoints = (T[IDX(Lf, i2 + 1, 0, k2 + 1)] +
                                         T[IDX(Lf, i2 + 1, 0, k2 - 1)] +
                                         T[IDX(Lf, i2 - 1, 0, k2 + 1)] +
                                         T[IDX(Lf, i2 - 1, 0, k2 - 1)]) * 0.0625;
                    Rc[IDX(Lc, i, 0, k)] = facePoints + vertPoints + T[IDX(Lf, i2, 0, k2)] * 0.25;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* prolong: d3:301-365, d2:239-271                                                             */
/* ------------------------------------------------------------------------------------------ */
static int isOdd(int x) { return x % 2; }

void mgo_prolong(mgo_solver *s) {
    const int pLevel = s->vLevel;
    s->vLevel -= 1;
    const level_dims *Lc = &s->lev[pLevel];
    const level_dims *Lf = &s->lev[s->vLevel];
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        const double *C = R->arr[MGO_PRESSURE][pLevel];
        double *Pf = R->arr[MGO_PRESSURE][s->vLevel];
        memset(Pf, 0, Lf->size * sizeof(double));
        if (s->prm.dim == 3) {
            for (int i = 0; i < Lf->n[0]; ++i) {
                int i2 = i / 2;
                for (int j = 0; j < Lf->n[1]; ++j) {
                    int j2 = j / 2;
                    for (int k = 0; k < Lf->n[2]; ++k) {
                        int k2 = k / 2;
                        double v;
                        if (isOdd(i)) {
                            if (isOdd(j)) {
                                if (isOdd(k))
                                    v = (C[IDX(Lc, i2, j2, k2)] +
                                         C[IDX(Lc, i2 + 1, j2, k2)] + C[IDX(Lc, i2, j2 + 1, k2)] + C[IDX(Lc, i2, j2, k2 + 1)] +
                                         C[IDX(Lc, i2 + 1, j2 + 1, k2)] + C[IDX(Lc, i2 + 1, j2, k2 + 1)] + C[IDX(Lc, i2, j2 + 1, k2 + 1)] +
                                         C[IDX(Lc, i2 + 1, j2 + 1, k2 + 1)]) / 8.0;
                                else
                                    v = (C[IDX(Lc, i2, j2, k2)] + C[IDX(Lc, i2, j2 + 1, k2)] +
                                         C[IDX(Lc, i2 + 1, j2, k2)] + C[IDX(Lc, i2 + 1, j2 + 1, k2)]) / 4.0;
                            } else {
                                if (isOdd(k))
                                    v = (C[IDX(Lc, i2, j2, k2)] + C[IDX(Lc, i2, j2, k2 + 1)] +
                                         C[IDX(Lc, i2 + 1, j2, k2)] + C[IDX(Lc, i2 + 1, j2, k2 + 1)]) / 4.0;
                                else
                                    v = (C[IDX(Lc, i2, j2, k2)] + C[IDX(Lc, i2 + 1, j2, k2)]) / 2.0;
                            }
                        } else {
                            if (isOdd(j)) {
                                if (isOdd(k))
                                    v = (C[IDX(Lc, i2, j2, k2)] + C[IDX(Lc, i2, j2, k2 + 1)] +
                                         C[IDX(Lc, i2, j2 + 1, k2)] + C[IDX(Lc, i2, j2 + 1, k2 + 1)]) / 4.0;
                                else
                                    v = (C[IDX(Lc, i2, j2, k2)] + C[IDX(Lc, i2, j2 + 1, k2)]) / 2.0;
                            } else {
                                if (isOdd(k))
                                    v = (C[IDX(Lc, i2, j2, k2)] + C[IDX(Lc, i2, j2, k2 + 1)]) / 2.0;
                                else
                                    v = C[IDX(Lc, i2, j2, k2)];
                            }
                        }
                        Pf[IDX(Lf, i, j, k)] = v;
                    }
                }
            }
        } else {
            for (int i = 0; i < Lf->n[0]; ++i) {
                int i2 = i / 2;
                for (int k = 0; k < Lf->n[2]; ++k) {
                    int k2 = k / 2;
                    double v;
                    if (isOdd(i)) {
                        if (isOdd(k))
                            v = (C[IDX(Lc, i2, 0, k2)] + C[IDX(Lc, i2, 0, k2 + 1)] +
                                 C[IDX(Lc, i2 + 1, 0, k2)] + C[IDX(Lc, i2 + 1, 0, k2 + 1)]) / 4.0;
                        else
                            v = (C[IDX(Lc, i2, 0, k2)] + C[IDX(Lc, i2 + 1, 0, k2)]) / 2.0;
                    } else {
                        if (isOdd(k))
                            v = (C[IDX(Lc, i2, 0, k2)] + C[IDX(Lc, i2, 0, k2 + 1)]) / 2.0;
                        else
                            v = C[IDX(Lc, i2, 0, k2)];
                    }
                    Pf[IDX(Lf, i, 0, k)] = v;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* computeError: d3:391-459, d2:274-338                                                        */
/* Sums run in the reference's loop order (i, j, k); MPI_SUM over ranks is taken in rank order.  */
/* ------------------------------------------------------------------------------------------ */
double mgo_compute_error(mgo_solver *s, int normOrder) {
    const level_dims *L = &s->lev[0];
    double numValGlo = 0.0, denValGlo = 0.0;
    int valCountGlo = 0;
    double residualVal = 0.0;
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        stencil_ctx c = make_ctx(s, R, 0);
        /* d3:411-416 uses hx(vLevel) etc.; vLevel is 0 whenever computeError is called (p:135) */
        if (s->vLevel != 0) {
            c.hx = s->coef[MGO_HX][s->vLevel]; c.hy = s->coef[MGO_HY][s->vLevel]; c.hz = s->coef[MGO_HZ][s->vLevel];
        }
        const double *P = R->arr[MGO_PRESSURE][0], *Rh = R->arr[MGO_RESIDUAL][0];
        double numValLoc = 0.0, denValLoc = 0.0, tempValue;
        int valCountLoc = 0;
        if (s->prm.dim == 3) {
            for (int i = 0; i < L->n[0]; ++i) for (int j = 0; j < L->n[1]; ++j) for (int k = 0; k < L->n[2]; ++k) {
                tempValue = fabs(lap3(&c, L, P, i, j, k) - Rh[IDX(L, i, j, k)]);
                switch (normOrder) {
                    case 1: if (tempValue > numValLoc) numValLoc = tempValue; break;
                    case 2: numValLoc += tempValue * tempValue;
                            denValLoc += Rh[IDX(L, i, j, k)] * Rh[IDX(L, i, j, k)];
                            valCountLoc += 1; break;
                }
            }
        } else {
            stencil_ctx c0 = make_ctx(s, R, 0);     /* d2:293-296 uses level-0 coefficients explicitly */
            for (int i = 0; i < L->n[0]; ++i) for (int k = 0; k < L->n[2]; ++k) {
                tempValue = fabs(lap2_err(&c0, L, P, i, k) - Rh[IDX(L, i, 0, k)]);
                switch (normOrder) {
                    case 1: if (tempValue > numValLoc) numValLoc = tempValue; break;
                    case 2: numValLoc += tempValue * tempValue;
                            denValLoc += Rh[IDX(L, i, 0, k)] * Rh[IDX(L, i, 0, k)];
                            valCountLoc += 1; break;
                }
            }
        }
        s->lups += (double)L->n[0] * L->n[1] * L->n[2];
        switch (normOrder) {
            case 1: {
                /* denValLoc = blitz::max(fabs(residualData(0))) over the whole padded array (d3:437) */
                double m = fabs(Rh[0]);
                for (size_t q = 1; q < L->size; q++) { double a = fabs(Rh[q]); if (a > m) m = a; }
                denValLoc = m;
                if (r == 0 || numValLoc > numValGlo) numValGlo = numValLoc;
                if (r == 0 || denValLoc > denValGlo) denValGlo = denValLoc;
                break;
            }
            case 2:
                numValGlo += numValLoc; denValGlo += denValLoc; valCountGlo += valCountLoc;
                break;
        }
    }
    switch (normOrder) {
        case 1:
            residualVal = denValGlo ? numValGlo / denValGlo : numValGlo;
            break;
        case 2:
            residualVal = denValGlo ? sqrt(numValGlo / valCountGlo) / sqrt(denValGlo / valCountGlo) : sqrt(numValGlo / valCountGlo);
            break;
    }
    return residualVal;
}

/* ------------------------------------------------------------------------------------------ */
/* vCycle: p:219-280                                                                           */
/* ------------------------------------------------------------------------------------------ */
static int g_last_coarse_iters;

void mgo_vcycle(mgo_solver *s) {
    const int D = s->prm.vcDepth;
    s->vLevel = 0;
    s->zeroBC = 0;
    mgo_smooth(s, s->prm.preSmooth);
    s->zeroBC = 1;
    for (int i = 0; i < D; i++) {
        mgo_compute_residual(s);
        for (int r = 0; r < s->nProc; r++)      /* p:252 smoothedPres = pressureData */
            memcpy(s->rk[r].arr[MGO_SMOOTHED][s->vLevel], s->rk[r].arr[MGO_PRESSURE][s->vLevel], s->lev[s->vLevel].size * sizeof(double));
        mgo_coarsen(s);
        for (int r = 0; r < s->nProc; r++)      /* p:258 */
            memset(s->rk[r].arr[MGO_PRESSURE][s->vLevel], 0, s->lev[s->vLevel].size * sizeof(double));
        if (s->vLevel == D) g_last_coarse_iters = mgo_solve_coarsest(s);
        else mgo_smooth(s, s->prm.preSmooth);
    }
    for (int i = 0; i < D; i++) {
        mgo_prolong(s);
        for (int r = 0; r < s->nProc; r++) {    /* p:271 */
            double *P = s->rk[r].arr[MGO_PRESSURE][s->vLevel];
            const double *S = s->rk[r].arr[MGO_SMOOTHED][s->vLevel];
            size_t n = s->lev[s->vLevel].size;
            for (size_t q = 0; q < n; q++) P[q] += S[q];
        }
        s->zeroBC = (s->vLevel == 0) ? 0 : 1;
        mgo_smooth(s, s->prm.postSmooth);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* mgSolve: p:100-205                                                                          */
/* ------------------------------------------------------------------------------------------ */
int mgo_mg_solve(mgo_solver *s, double *const *p_inout, const double *const *rhs,
                 double *res_hist, int *coarse_iters, double *max_dev) {
    const int D = s->prm.vcDepth;
    const level_dims *L0 = &s->lev[0];
    double prevResidual = 0.0;
    int cycles = 0;
    s->vLevel = 0;

    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        for (int l = 0; l <= D; l++) {                      /* p:107-111 */
            memset(R->arr[MGO_PRESSURE][l], 0, s->lev[l].size * sizeof(double));
            memset(R->arr[MGO_RESIDUAL][l], 0, s->lev[l].size * sizeof(double));
            memset(R->arr[MGO_SMOOTHED][l], 0, s->lev[l].size * sizeof(double));
        }
        for (int i = 0; i < L0->n[0]; i++) for (int j = 0; j < L0->n[1]; j++) for (int k = 0; k < L0->n[2]; k++) {
            R->arr[MGO_RESIDUAL][0][IDX(L0, i, j, k)] = rhs[r][IDX(L0, i, j, k)];           /* p:114 */
            R->arr[MGO_PRESSURE][0][IDX(L0, i, j, k)] = p_inout[r][IDX(L0, i, j, k)];       /* p:115 */
        }
    }
    mgo_update_pads(s, MGO_RESIDUAL);                       /* p:117 */
    mgo_update_pads(s, MGO_PRESSURE);                       /* p:118 */

    if (!s->prm.test_poisson) {                             /* p:120-129 */
        double globalAvg = 0.0;
        for (int r = 0; r < s->nProc; r++) {
            const double *Rh = s->rk[r].arr[MGO_RESIDUAL][0];
            double sum = 0.0;                               /* blitz::mean: sequential sum in storage order / count */
            if (s->exactMean) sum = exact_sum(Rh, L0->size);
            else for (size_t q = 0; q < L0->size; q++) sum += Rh[q];
            double localMean = sum / (double)L0->size;
            globalAvg += localMean;                         /* MPI_SUM in rank order */
        }
        globalAvg /= s->nProc;
        s->lastGlobalAvg = globalAvg;
        for (int r = 0; r < s->nProc; r++) {
            double *Rh = s->rk[r].arr[MGO_RESIDUAL][0];
            for (size_t q = 0; q < L0->size; q++) Rh[q] -= globalAvg;
        }
    }

    for (int i = 0; i < s->prm.vcCount; i++) {              /* p:132-148 */
        g_last_coarse_iters = 0;
        mgo_vcycle(s);
        double mgResidual = mgo_compute_error(s, s->prm.resType);
        if (res_hist) res_hist[i] = mgResidual;
        if (coarse_iters) coarse_iters[i] = g_last_coarse_iters;
        cycles = i + 1;
        if (!s->prm.test_poisson) {
            if (fabs(prevResidual - mgResidual) < s->prm.mgTolerance) break;
            prevResidual = mgResidual;
        }
    }

    for (int r = 0; r < s->nProc; r++)                      /* p:151 whole padded box */
        memcpy(p_inout[r], s->rk[r].arr[MGO_PRESSURE][0], L0->size * sizeof(double));

    if (s->prm.test_poisson && max_dev) {                   /* p:153-204 */
        double gloMax = 0.0;
        const double hx0 = s->coef[MGO_HX][0], hy0 = s->coef[MGO_HY][0], hz0 = s->coef[MGO_HZ][0];
        const int ubx = L0->n[0] - 1, uby = L0->n[1] - 1, ubz = L0->n[2] - 1;
        for (int r = 0; r < s->nProc; r++) {
            rank_state *R = &s->rk[r];
            const double *P = R->arr[MGO_PRESSURE][0];
            int halfIndX = ubx * s->prm.npX / 2;
            int halfIndY = uby * s->prm.npY / 2;
            double locMax = -HUGE_VAL;
            for (int i = 0; i <= ubx; ++i) {
                double xDist = hx0 * (R->xRank * ubx + i - halfIndX);
                for (int j = 0; j <= uby; ++j) {
                    double yDist = hy0 * (R->yRank * uby + j - halfIndY);
                    for (int k = 0; k <= ubz; ++k) {
                        double zDist = hz0 * (k - ubz / 2);
                        double pa = (s->prm.dim == 3) ? (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0
                                                      : (xDist * xDist + zDist * zDist) / 4.0;
                        double d = fabs(pa - P[IDX(L0, i, j, k)]);
                        if (d > locMax) locMax = d;
                    }
                }
            }
            if (locMax > gloMax) gloMax = locMax;
        }
        *max_dev = gloMax;
    }
    return cycles;
}

/* ------------------------------------------------------------------------------------------ */
/* testProlong: d3:689-719, d2:481-507                                                         */
/* ------------------------------------------------------------------------------------------ */
double mgo_test_prolong(mgo_solver *s) {
    double worst = 0.0;
    const int planar = (s->prm.dim == 2);
    s->vLevel = 0;
    for (int r = 0; r < s->nProc; r++) {
        rank_state *R = &s->rk[r];
        for (int l = 0; l <= 1; l++) {
            const level_dims *L = &s->lev[l];
            double *A = R->arr[l == 0 ? MGO_RESIDUAL : MGO_PRESSURE][l];
            memset(A, 0, L->size * sizeof(double));
            for (int i = 0; i < L->n[0]; ++i) for (int j = 0; j < L->n[1]; ++j) for (int k = 0; k < L->n[2]; ++k)
                A[IDX(L, i, j, k)] = planar ? (R->rank + 1) * 100 + i * 10 + k : (R->rank + 1) * 1000 + i * 100 + j * 10 + k;
        }
    }
    s->vLevel = 1;
    mgo_prolong(s);
    for (int r = 0; r < s->nProc; r++) {
        const level_dims *L = &s->lev[0];
        const double *P = s->rk[r].arr[MGO_PRESSURE][0], *E = s->rk[r].arr[MGO_RESIDUAL][0];
        for (size_t q = 0; q < L->size; q++) { double d = fabs(P[q] - E[q]); if (d > worst) worst = d; }
    }
    return worst;
}


