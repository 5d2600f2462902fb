/*
 * mg_oracle.h -- CPU parity oracle for the SARAS pressure-Poisson multigrid path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under saras_b200/ may include, link or load
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported baseline.
 *
 * The oracle is a plain-C restatement (explicit loops, no Blitz, no MPI) of
 *   /root/reference/lib/poisson/poisson.cc, poisson_d3.cc, poisson_d2.cc
 *   /root/reference/lib/grid/grid.cc      (metric arrays only)
 *   /root/reference/lib/parallel/parallel.cc (rank topology only)
 *   /root/reference/src/solvers/hydro/hydro.cc:174-203 (checkPeriodic)
 * Every function in mg_oracle.c cites the file:line it follows.
 *
 * All npX*npY MPI ranks of the reference are emulated inside one process: each
 * bulk-synchronous operator is run rank by rank, and a halo exchange is a
 * global step (x faces for every rank, then y faces for every rank; see
 * mgo_update_pads for why that order is the deterministic reading of the
 * reference's racy Irecv/Send sequence).
 *
 * Parity pin: oracle/refbuild compiles the reference's own poisson*.cc, grid.cc
 * and parallel.cc UNMODIFIED against stand-in Blitz/MPI/yaml headers and
 * tests/test_oracle_vs_ref.py checks this restatement against it bit for bit;
 * the outputs are also committed under tests/golden/.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int dim;            /* 3, or 2 for a -DPLANAR build (poisson.h:156 multigrid_d2) */
    int test_poisson;   /* semantics of -DTEST_POISSON (poisson.cc:101,120,137,144,153) */
    int xInd, yInd, zInd;       /* parser.h:71, grid sizes 2^Ind + 1 */
    int npX, npY;               /* parser.h:69 */
    int xPer, yPer, zPer;       /* parser.cc:406-414 */
    int xGrid, yGrid, zGrid;    /* parser.cc:361-391: 0 uniform, 2 tanh ("D") */
    double betaX, betaY, betaZ;
    double Lx, Ly, Lz;
    int vcDepth, vcCount, preSmooth, postSmooth, gsSmooth, resType;
    double mgTolerance;
    int wall_null;      /* 1: hydro::checkPeriodic ran (faceRanks at walls = PROC_NULL, hydro.cc:183-201)
                           0: faceRanks fully periodic (state seen by the -DTEST_RUN self tests) */
} mgo_params;

typedef struct mgo_solver mgo_solver;

/* arrays of a level, in the reference's names (poisson.h:70-73) */
enum { MGO_PRESSURE = 0, MGO_RESIDUAL = 1, MGO_TMP = 2, MGO_SMOOTHED = 3 };
/* metric arrays (poisson.h:91-93) */
enum { MGO_XIXX = 0, MGO_XIX2 = 1, MGO_ETYY = 2, MGO_ETY2 = 3, MGO_ZTZZ = 4, MGO_ZTZ2 = 5 };
/* scalar coefficients per level (poisson.h:87-89) */
enum { MGO_HX = 0, MGO_HY, MGO_HZ, MGO_HX2, MGO_HZ2, MGO_HZHX, MGO_HXHY, MGO_HYHZ, MGO_HXHYHZ, MGO_NCOEF };

mgo_solver *mgo_create(const mgo_params *p);
void mgo_destroy(mgo_solver *s);

int mgo_nranks(const mgo_solver *s);
/* core points n[3] and padded extents F[3] of `level` (identical on every rank) */
void mgo_level_dims(const mgo_solver *s, int level, int n[3], int F[3]);
/* rank topology as the reference computes it: out = {xRank, yRank, face0..face3}; -1 = MPI_PROC_NULL */
void mgo_rank_info(const mgo_solver *s, int rank, int out[6]);
/* padded metric array (length F_d of that level, pads = 0), copied to out */
void mgo_metric(const mgo_solver *s, int rank, int level, int which, double *out);
double mgo_coef(const mgo_solver *s, int level, int which);
/* grid scalars: out = {dXi, dEt, dZt} */
void mgo_grid_spacing(const mgo_solver *s, double out[3]);
/* staggered coordinates x/y/zStaggr of rank (length F_d(0), index base -1) */
void mgo_staggr_coords(const mgo_solver *s, int rank, int dir, double *out);
/* TEST_POISSON Dirichlet planes (poisson_d3.cc:513-572 / d2:367-400); which = 0 x, 1 y, 2 z; padded, row-major */
void mgo_wall(const mgo_solver *s, int rank, int which, double *out);
int mgo_max_count(const mgo_solver *s);

/* direct access to a level array of a rank (padded, z fastest, base index -1) */
double *mgo_array(mgo_solver *s, int rank, int level, int which);

/* the reference's protected operators (poisson.h:112-123), acting on all ranks */
void mgo_set_level(mgo_solver *s, int vLevel);
int  mgo_get_level(const mgo_solver *s);
void mgo_set_zero_bc(mgo_solver *s, int zeroBC);
void mgo_update_pads(mgo_solver *s, int which);
void mgo_impose_bc(mgo_solver *s);
void mgo_smooth(mgo_solver *s, int count);
void mgo_compute_residual(mgo_solver *s);
void mgo_coarsen(mgo_solver *s);
void mgo_prolong(mgo_solver *s);
int  mgo_solve_coarsest(mgo_solver *s);          /* returns iterCount */
double mgo_compute_error(mgo_solver *s, int normOrder);
void mgo_vcycle(mgo_solver *s);

/* poisson::mgSolve (poisson.cc:100-205).  p_inout[r], rhs[r]: padded level-0 arrays of rank r.
 * res_hist (>= vcCount doubles) receives the value the reference prints after every V-cycle,
 * coarse_iters (>= vcCount ints, may be NULL) the iterCount of solve() in that cycle.
 * max_dev (may be NULL): TEST_POISSON "Maximum absolute deviation from analytic solution".
 * Returns the number of V-cycles executed. */
int mgo_mg_solve(mgo_solver *s, double *const *p_inout, const double *const *rhs,
                 double *res_hist, int *coarse_iters, double *max_dev);

/* testProlong of the reference (poisson_d3.cc:689-719, poisson_d2.cc:481-507); max over ranks.  testTransfer and
 * testPeriodic (d3:722-857) have no oracle twin: their expected values are analytic (the neighbour's index pattern, the
 * periodic continuation of sin*cos*cos) and the GPU self tests compare against those directly. */
double mgo_test_prolong(mgo_solver *s);

/* LUP counters since creation/reset: smoother + residual + error-norm stencil evaluations */
double mgo_lup_count(const mgo_solver *s);
void mgo_lup_reset(mgo_solver *s);
/* globalAvg subtracted from the RHS by the last mgo_mg_solve (poisson.cc:122-128) */
double mgo_last_global_avg(const mgo_solver *s);
/* TEST MODE, not the reference's arithmetic: form the local sums of that mean correctly rounded (exact summation)
 * instead of with the reference's one sequential fp64 sum; see exact_sum in mg_oracle.c */
void mgo_set_exact_mean(mgo_solver *s, int on);

#ifdef __cplusplus
}
#endif
#endif
