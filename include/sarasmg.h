/*
 * sarasmg.h -- C ABI of the B200-native multigrid pressure-Poisson solver.
 *
 * This is the drop-in boundary for SARAS's lib/poisson (reference file:line cited per
 * entry point, paths relative to the SARAS source tree).  Plain C types only: no Blitz,
 * MPI, torch or CUDA types cross it.  The host side (saras_b200/host/poisson.{h,cc}, the
 * replacement for lib/poisson/poisson.h + poisson*.cc) fills a sarasmg_desc from the
 * reference's `grid`, `parser` and `parallel` objects and forwards `poisson::mgSolve`.
 *
 * Array conventions (identical to the reference's Blitz arrays, lib/poisson/poisson.cc:293-316):
 *   a level array has padded extents F_d = n_d + 2, index base -1 in every direction,
 *   row-major with z fastest:  off(i,j,k) = ((i+1)*F_y + (j+1))*F_z + (k+1).
 *   2-D (PLANAR) arrays are F_x x 3 x F_z with only j = 0 live (poisson.cc:339-343).
 *   Host pointers passed here always use this compact layout; the device layout is private.
 *
 * All functions return 0 on success, non-zero on failure; sarasmg_last_error() describes the
 * failure.  There is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef SARASMG_H
#define SARASMG_H

#ifdef __cplusplus
extern "C" {
#endif

#define SARASMG_PROC_NULL (-1)      /* stands for MPI_PROC_NULL in face_ranks */
#define SARASMG_UNIQUE_ID_BYTES 128 /* sizeof(ncclUniqueId) */

/* desc.flags */
#define SARASMG_FLAG_TIME_RUN   1u  /* keep the -DTIME_RUN compute/transfer split (poisson.h:60-63, poisson_d3.cc:115-186) */
#define SARASMG_FLAG_PROFILE    2u  /* bracket every operator launch with CUDA events (sarasmg_profile_query) */

typedef struct sarasmg_solver_s *sarasmg_handle;

/* arrays of a level, in the reference's names (poisson.h:70-73).  smoothedPres has no device
 * copy: the solver keeps pressureData in place across the coarse-grid visit. */
enum { SARASMG_PRESSURE = 0, SARASMG_RESIDUAL = 1, SARASMG_TMP = 2 };

typedef struct {
    int dim;                /* 3 = multigrid_d3, 2 = multigrid_d2 (-DPLANAR)                       poisson.h:156-219 */
    int test_poisson;       /* semantics of -DTEST_POISSON                                          poisson.cc:101,120,137,144; poisson_d3.cc:579-666 */
    int vc_depth;           /* parser::vcDepth                                                      parser.cc:225 */
    int vc_count;           /* parser::vcCount */
    int pre_smooth;         /* parser::preSmooth */
    int post_smooth;        /* parser::postSmooth */
    int gs_smooth;          /* parser::gsSmooth: 0 Jacobi, 1 lexicographic Gauss-Seidel              poisson_d3.cc:135 */
    int res_type;           /* parser::resType 0/1/2                                                poisson_d3.cc:418-455 */
    int print_residual;     /* parser::printResidual                                                poisson.cc:136 */
    double mg_tolerance;    /* parser::mgTolerance                                                  poisson.cc:145, poisson_d3.cc:235 */
    int x_per, y_per, z_per;/* parser::xPer/yPer/zPer                                               poisson_d3.cc:578,610,642 */

    /* parallel object (lib/parallel/parallel.h:51-122), read at the first mgSolve so that the
     * MPI_PROC_NULL edits of hydro::checkPeriodic (hydro.cc:183-201) are seen */
    int rank, n_proc, npx, npy, x_rank, y_rank;
    int face_ranks[4];      /* left, right, front, back; SARASMG_PROC_NULL at walls */

    int n0[3];              /* level-0 core points per direction of this rank = stagCore(0).ubound()+1   poisson.cc:336-357 */

    /* per-level scalars, vc_depth+1 doubles each, computed on the HOST with the reference's
     * pow() expressions (poisson.cc:419-438).  3-D uses hx hy hz hxhy hyhz hzhx hxhyhz, 2-D hx hz hx2 hz2 hzhx. */
    const double *hx, *hy, *hz, *hx2, *hz2, *hzhx, *hxhy, *hyhz, *hxhyhz;

    /* per-level 1-D metric arrays as produced by poisson::copyStaggrDerivs (poisson.cc:453-496):
     * padded (F_d(l) doubles, pads = 0), levels concatenated 0..vc_depth.  etyy/ety2 ignored in 2-D. */
    const double *xixx, *xix2, *etyy, *ety2, *ztzz, *ztz2;

    /* TEST_POISSON Dirichlet ghost planes of level 0 (poisson_d3.cc:513-572, poisson_d2.cc:367-400):
     * 3-D: x_wall F_y*F_z, y_wall F_x*F_z, z_wall F_x*F_y;  2-D: x_wall F_z, z_wall F_x.  NULL otherwise. */
    const double *x_wall, *y_wall, *z_wall;

    int device;             /* CUDA device ordinal of this rank */
    const void *nccl_id;    /* SARASMG_UNIQUE_ID_BYTES bytes shared by all ranks (sarasmg_get_unique_id on rank 0,
                               broadcast by the host: MPI_Bcast in SARAS); NULL when n_proc == 1 */
    unsigned flags;
} sarasmg_desc;

const char *sarasmg_last_error(void);
int sarasmg_version(void);

/* number of usable CUDA devices (0 and an error when there is none) */
int sarasmg_device_count(int *count);
/* fills out[SARASMG_UNIQUE_ID_BYTES] with a fresh ncclUniqueId */
int sarasmg_get_unique_id(void *out);

/* multigrid_d2/d3 constructor: setStagBounds, initializeArrays, createMGSubArrays (poisson_d3.cc:64-87).
 * Allocates the device-resident level hierarchy, which persists until sarasmg_destroy. Collective over ranks
 * when n_proc > 1 (creates the NCCL communicator). */
int sarasmg_create(const sarasmg_desc *desc, sarasmg_handle *out);
int sarasmg_destroy(sarasmg_handle h);

/* poisson::mgSolve(plainsf &inFn, const plainsf &rhs)                                     poisson.cc:100-205
 * p_inout / rhs: HOST level-0 padded arrays (inFn.F.dataFirst(), rhs.F.dataFirst()).  Only the core of both is
 * read (poisson.cc:114-115); the whole padded box of p_inout is overwritten (poisson.cc:151).
 * cycles_done: V-cycles executed; residual_history[>= vc_count]: the value printed after every cycle
 * (poisson.cc:138-140); coarse_iters[>= vc_count]: iterCount of solve() per cycle (poisson_d3.cc:239);
 * any of the three may be NULL.  Collective over ranks. Prints the reference's lines on rank 0 when print_residual. */
int sarasmg_solve(sarasmg_handle h, double *p_inout, const double *rhs,
                  int *cycles_done, double *residual_history, int *coarse_iters);

/* The same solve split at the PCIe boundary, for callers that keep Pp / mgRHS device-resident:
 * sarasmg_upload copies both host arrays into device input buffers; sarasmg_solve_resident runs everything
 * mgSolve does between the two copies (zero hierarchy, core load, pad sync, mean removal, V-cycles) from
 * those buffers; sarasmg_download copies pressureData(0) back. */
int sarasmg_upload(sarasmg_handle h, const double *p_in, const double *rhs);
int sarasmg_solve_resident(sarasmg_handle h, int *cycles_done, double *residual_history, int *coarse_iters);
int sarasmg_download(sarasmg_handle h, double *p_out);
/* ---- the projection step around mgSolve, device-resident ------------------------------------------------
 * hydro_d3::timeAdvance brackets the solve with two bandwidth-bound field operations (hydro_d3.cc:326-327, 347-362):
 *     V.divergence(mgRHS, P);  mgRHS *= 1.0/dt;                      lib/field/vfield.cc:299-320
 *     mgSolver.mgSolve(Pp, mgRHS);
 *     Pp.syncData();  P += Pp;                                       lib/parallel/mpidata.cc:278-306, lib/field/sfield.cc:197-201
 *     Pp.gradient(pressureGradient, V);  pressureGradient *= dt;  V -= pressureGradient;      lib/field/plainsf.h:92-98
 * With host fields every solve pays two PCIe crossings of Pp / mgRHS.  These entries keep V, P, Pp and mgRHS on the
 * device: the caller uploads the velocity once (or after its own host-side stages), then each sarasmg_proj_step runs
 * divergence -> mgSolve -> correction without touching the host; Pp stays resident as the next step's initial guess.
 * Field layouts are the reference's Blitz arrays (base -1, z fastest; lib/field/field.cc:66-105): Vx (nx+1, ny+2, nz+2),
 * Vy (nx+2, ny+1, nz+2), Vz (nx+2, ny+2, nz+1), P / Pp / mgRHS (nx+2, ny+2, nz+2), nx.. = n0[] of the descriptor.
 * Several ranks: Pp.syncData() exchanges the four faces with face_ranks (core rows only) and the four edge
 * columns with the diagonal ranks (always the periodic wrap, parallel.cc:124-129) in one grouped ncclSend/ncclRecv on the
 * solver's communicator.  multigrid_d3 only. */
typedef struct {
    const double *xi_x_staggr, *et_y_staggr, *zt_z_staggr;   /* grid::xi_xStaggr ..: n_d + 2 doubles, first entry is index -1 */
    const double *xi_x_colloc, *et_y_colloc, *zt_z_colloc;   /* grid::xi_xColloc ..: n_d + 1 doubles, first entry is index -1 */
    double d_xi, d_et, d_zt;                                 /* grid::dXi, dEt, dZt */
} sarasmg_proj_desc;
enum { SARASMG_PROJ_VX = 0, SARASMG_PROJ_VY = 1, SARASMG_PROJ_VZ = 2, SARASMG_PROJ_P = 3, SARASMG_PROJ_PP = 4, SARASMG_PROJ_RHS = 5 };
int sarasmg_proj_setup(sarasmg_handle h, const sarasmg_proj_desc *desc);
int sarasmg_proj_upload(sarasmg_handle h, int which, const double *host);      /* VX, VY, VZ, P, PP */
int sarasmg_proj_download(sarasmg_handle h, int which, double *host);          /* ... and RHS */
int sarasmg_proj_divergence(sarasmg_handle h, double inv_dt);                  /* mgRHS = div(V) * inv_dt       vfield.cc:299-320, hydro_d3.cc:327 */
int sarasmg_proj_correct(sarasmg_handle h, double dt);                         /* Pp = solution; syncData; P += Pp; V -= dt grad(Pp)   hydro_d3.cc:347-362 */
/* divergence -> mgSolve -> correction in one call; history arguments as in sarasmg_solve */
int sarasmg_proj_step(sarasmg_handle h, double dt, int *cycles_done, double *residual_history, int *coarse_iters);

/* ---- the implicit velocity solves of the time stepper, device-resident -------------------------------------
 * hydro_d3::solveVx / solveVy / solveVz (src/solvers/hydro/hydro_d3.cc:370-533): Jacobi iterations of
 * (1 - dt/(2 Re) Lap) V = nseRHS on the velocity component held by the projection entries above (sarasmg_proj_upload VX / VY /
 * VZ), with hydro::imposeUBCs / VBCs / WBCs (hydro.cc:386-445) after every sweep -- syncData, then the wall objects of
 * hydro::initVBC in the order left, right, front, back, top, bottom -- until max |residual| < cn_tolerance.  The result
 * replaces the component on the device (the next sarasmg_proj_divergence reads it).  One rank; multigrid_d3 only. */
enum { SARASMG_BC_NONE = 0, SARASMG_BC_DIRICHLET_CC, SARASMG_BC_DIRICHLET_FC, SARASMG_BC_NEUMANN, SARASMG_BC_PERIODIC_CC, SARASMG_BC_PERIODIC_FC };
typedef struct {
    double dt, Re, cn_tolerance;                        /* time step, inputParams.Re, inputParams.cnTolerance */
    double hx, hy, hz;                                  /* hydro::setCoefficients, hydro.cc:223-241 (mesh.dXi, dEt, dZt ...) */
    double hy2hz2, hz2hx2, hx2hy2, hx2hy2hz2;           /* ... and the host's pow() products */
    const double *xix2_colloc, *ety2_colloc, *ztz2_colloc;      /* grid::xix2Colloc ..: n_d + 1 doubles, first entry is index -1 */
    const double *xix2_staggr, *ety2_staggr, *ztz2_staggr;      /* grid::xix2Staggr ..: n_d + 2 doubles */
    int bc_type[3][6];                                  /* [component][wall]: the boundary class of hydro::initVBC (SARASMG_BC_*), NONE on periodic x / y walls */
    double bc_value[3][6];
    long long max_iterations;                           /* hydro.cc:74: product of the collocated core sizes */
} sarasmg_impl_desc;
int sarasmg_impl_setup(sarasmg_handle h, const sarasmg_impl_desc *desc);
int sarasmg_impl_upload_rhs(sarasmg_handle h, int component, const double *host);     /* nseRHS.Vx / Vy / Vz, layout of the component */
int sarasmg_impl_bcs(sarasmg_handle h, int component);                                /* imposeUBCs / VBCs / WBCs alone */
/* sweeps: Jacobi sweeps performed (iterCount + 1 of the reference); fails when max_iterations is exceeded (the reference aborts) */
int sarasmg_impl_solve(sarasmg_handle h, int component, int *sweeps, double *max_error);

/* TEST_POISSON epilogue (poisson.cc:153-204): max |analytic - pressureData(0)| over the core, all ranks */
int sarasmg_analytic_deviation(sarasmg_handle h, double *max_dev);

/* ---- the reference's protected operators, exposed for operator-level parity tests (poisson.h:112-123) ---- */
int sarasmg_level_dims(sarasmg_handle h, int level, int n[3], int F[3]);
int sarasmg_level_set(sarasmg_handle h, int level, int which, const double *host);   /* compact padded host array -> device */
int sarasmg_level_get(sarasmg_handle h, int level, int which, double *host);
int sarasmg_set_level(sarasmg_handle h, int v_level);                                 /* poisson::vLevel */
int sarasmg_set_zero_bc(sarasmg_handle h, int zero_bc);                               /* poisson::zeroBC */
int sarasmg_op_update_pads(sarasmg_handle h, int which);     /* updatePads        poisson_d3.cc:671-686, poisson_d2.cc:467-478 */
int sarasmg_op_impose_bc(sarasmg_handle h);                  /* imposeBC          poisson_d3.cc:575-668, poisson_d2.cc:403-464 */
int sarasmg_op_smooth(sarasmg_handle h, int count);          /* smooth            poisson_d3.cc:114-187, poisson_d2.cc:109-147 */
int sarasmg_op_compute_residual(sarasmg_handle h);           /* computeResidual   poisson_d3.cc:90-111,  poisson_d2.cc:90-106 */
int sarasmg_op_coarsen(sarasmg_handle h);                    /* coarsen           poisson_d3.cc:250-298, poisson_d2.cc:202-236 */
int sarasmg_op_prolong(sarasmg_handle h);                    /* prolong + "+= smoothedPres"  poisson_d3.cc:301-365, poisson.cc:271 */
int sarasmg_op_solve_coarsest(sarasmg_handle h, int *iters); /* solve             poisson_d3.cc:190-247, poisson_d2.cc:150-199 */
int sarasmg_op_compute_error(sarasmg_handle h, int norm_order, double *value); /* computeError poisson_d3.cc:391-459 */
int sarasmg_op_vcycle(sarasmg_handle h, int *coarse_iters);  /* vCycle            poisson.cc:219-280 */

/* ---- self tests (poisson.h:130-132; poisson_d3.cc:689-857).  Each returns the max abs deviation. ---- */
int sarasmg_selftest_prolong(sarasmg_handle h, double *max_err);
int sarasmg_selftest_transfer(sarasmg_handle h, double *max_err);
int sarasmg_selftest_periodic(sarasmg_handle h, const double *x_staggr, const double *y_staggr, const double *z_staggr,
                              double x_len, double y_len, double z_len, double *max_err);

/* ---- instrumentation ---- */
/* -DTIME_RUN contract (poisson.cc:645-652): seconds spent in smoother sweeps vs imposeBC inside smooth() */
int sarasmg_timers(sarasmg_handle h, double *smooth_compute_s, double *smooth_transfer_s);
/* number of CUDA kernels this library launched for the handle since creation / last reset */
long long sarasmg_kernel_launches(sarasmg_handle h);
int sarasmg_reset_counters(sarasmg_handle h);
/* operator classes for profiling */
enum { SARASMG_K_SMOOTH = 0, SARASMG_K_RESIDUAL, SARASMG_K_COARSEN, SARASMG_K_PROLONG, SARASMG_K_ERROR,
       SARASMG_K_BC, SARASMG_K_COARSE_SOLVE, SARASMG_K_HALO, SARASMG_K_MISC,
       SARASMG_K_FUSED_RC /* computeResidual + coarsen as one pass (largest levels) */, SARASMG_K_COUNT };
/* with SARASMG_FLAG_PROFILE: accumulated CUDA-event time (ms) and launch count of operator class `kind` on `level`
 * since the last sarasmg_reset_counters.  Synchronises the stream. */
int sarasmg_profile_query(sarasmg_handle h, int kind, int level, double *total_ms, long long *launches);
/* device-side time of the last sarasmg_solve_resident / sarasmg_solve V-cycle loop (CUDA events), ms */
int sarasmg_last_solve_ms(sarasmg_handle h, double *ms);
/* with SARASMG_FLAG_PROFILE: only operators on levels <= max_level are bracketed with events (default: all levels).
 * Lets a benchmark time the bandwidth-bound fine levels live without adding host work to the latency-bound coarse ones. */
int sarasmg_set_profile_max_level(sarasmg_handle h, int max_level);
/* user timing events on the SOLVER's stream (the stream every kernel of the handle is launched on): record slot 0..7,
 * then elapsed(a, b) synchronises on b and returns the device time between the two records in ms */
int sarasmg_event_record(sarasmg_handle h, int slot);
int sarasmg_event_elapsed_ms(sarasmg_handle h, int slot_a, int slot_b, double *ms);
/* page-lock a caller-owned host range (the Blitz storage of Pp / mgRHS, which SARAS allocates once and reuses every
 * time step) so that the copies of sarasmg_solve run at PCIe speed.  Nothing is ever pinned implicitly: the library cannot
 * know the lifetime of a caller's buffer.  LIFETIME: the range must stay allocated until sarasmg_host_unregister(ptr) or
 * sarasmg_destroy(h) -- freeing registered memory is undefined behaviour (CUDA).  The C++ drop-in registers the contiguous
 * storage of the plainsf objects it is handed (never its own temporary staging copies) and unregisters them in ~poisson,
 * which runs before the fields die in every SARAS solver (hydro.h: Pp / mgRHS are base-class members, mgSolver a
 * derived-class member). */
int sarasmg_host_register(sarasmg_handle h, const void *ptr, unsigned long long bytes);
int sarasmg_host_unregister(sarasmg_handle h, const void *ptr);
/* mgSolve subtracts globalAvg = mean of the padded RHS (poisson.cc:122-128).  The reference forms that mean with ONE
 * sequential fp64 sum, whose rounding error (~sqrt(n) ulp of the sum) depends on the traversal order of blitz::mean and is
 * not reproduced by any parallel reduction; the singular Neumann / periodic solve amplifies it ~100x in the field.  The
 * device forms the CORRECTLY ROUNDED sum (compensated summation), i.e. the value the reference's sum approximates.
 * TEST HOOK: with enable != 0 the solver subtracts `value` instead, so a test can feed the oracle's mean and then require
 * bit-exact fields in production mode as well.  sarasmg_last_rhs_mean returns the globalAvg of the last solve. */
int sarasmg_debug_set_rhs_mean(sarasmg_handle h, int enable, double value);
int sarasmg_last_rhs_mean(sarasmg_handle h, double *value);
/* parser::vcCount of the handle (the reference reads it once from the YAML; a benchmark uses this to run the same hierarchy
 * to stagnation for the time-to-tolerance figure).  History buffers passed to the solve calls must hold `count` entries. */
int sarasmg_set_vc_count(sarasmg_handle h, int count);
/* parser::gsSmooth of the handle (0 Jacobi, 1 lexicographic Gauss-Seidel): lets a benchmark time both smoothers on one
 * resident hierarchy.  Takes effect at the next solve / operator call. */
int sarasmg_set_smoother(sarasmg_handle h, int gs_smooth);
/* blocks until all work of the handle has finished */
int sarasmg_synchronize(sarasmg_handle h);

#ifdef __cplusplus
}
#endif
#endif /* SARASMG_H */
