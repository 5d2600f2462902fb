/*
 * sarasmg.cu -- device-resident multigrid hierarchy, V-cycle driver and the extern "C" layer
 * declared in include/sarasmg.h.  Control flow follows SARAS's lib/poisson/poisson.cc (mgSolve
 * :100-205, vCycle :219-280) and poisson_d3.cc / poisson_d2.cc statement by statement; the
 * arithmetic lives in mg_kernels.cuh.  There is no CPU fallback.
 *
 * Differences from the reference's data handling that do not change any result:
 *   - smoothedPres (poisson.h:73) is not stored.  pressureData(l) is dead between p:252 and p:271
 *     (it is zeroed by prolong() at d3:308 before being read again), so the device keeps the
 *     pre-coarsening iterate in place and k_prolong_add adds the interpolated correction to it.
 *   - tmpDataArray(l) = 0 (d3:91, d3:118) is reduced to zeroing its pads: the core is overwritten
 *     before it is read, and in smooth() every pad plane is rewritten by imposeBC.
 *   - swap(tmpDataArray, pressureData) (d3:168) swaps every level at once in the reference; all the
 *     other levels are dead or re-initialised before use, so a per-level pointer swap is equivalent.
 */
#include "../../include/sarasmg.h"
#include "mg_kernels.cuh"
#include "gs_skew.cuh"
#include "gs_strip.cuh"
#include "mg_tma.cuh"
#include "mg_fused.cuh"
#include "proj_kernels.cuh"
#include "impl_kernels.cuh"

#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;

static int set_err(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return set_err("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)
#define NC(call)                                                                                     \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess) return set_err("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(r_)); \
    } while (0)
#define TRY(call)                   \
    do {                            \
        int rc_ = (call);           \
        if (rc_) return rc_;        \
    } while (0)

struct Level {
    LevelView v;
    bool x_uniform;         /* xix2 has one value over the core: the smoother's diagonal does not change along an x-march */
    bool z_uniform;         /* ztz2 has one value over the core: the diagonal does not change along a z column */
    bool zz_uniform;        /* ztz2 and ztzz each have one value over the core */
    GsPlan gs;              /* skewed Gauss-Seidel sweep: task table + progress counters (device) */
    int4 *gs_tasks;
    int *gs_sync;
    size_t gs_smem;
    GsStripPlan gss;        /* warp-strip Gauss-Seidel sweep of the large 3-D levels (gs_strip.cuh) */
    int2 *gss_tasks;
    int *gss_sync;
    double *gss_rden, *gss_xpair;
    size_t gss_smem;
    bool gss_ok;
    double *buf[3];         /* SARASMG_PRESSURE, SARASMG_RESIDUAL, SARASMG_TMP: carved out of the solver's one slab */
    size_t off[3];          /* their offsets (doubles) in the slab: the same on every rank, so a neighbour's copy of an array
                               is its mapped slab + off; swapped together with buf[] */
    CUtensorMap tmapP[3];   /* TMA descriptors of the three arrays for the pressure box (with halo ring) ... */
    CUtensorMap tmapR[3];   /* ... and for the right-hand-side box; swapped together with buf[] */
    CUtensorMap tmapP16[3], tmapR16[3];     /* the same for the 16-row tiles of the fused residual + restriction kernel (mg_fused.cuh) */
    bool tma;               /* level is large enough for k_stencil_tma and the descriptors exist */
    double *met[6];         /* device metric arrays (padded, base -1) */
    size_t elems;           /* allocated doubles per array */
};

struct ProfEntry { int kind, level, bucket; cudaEvent_t a, b; };

/* one captured V-cycle: valid for one assignment of the level buffers (Jacobi ping-pong) and one profiling setup */
struct GraphSlot {
    const double *key;              /* lev[0].buf[SARASMG_PRESSURE] at entry */
    unsigned prof_key;
    cudaGraphExec_t exec;
    std::vector<ProfEntry> prof;    /* external event pairs recorded by the graph, re-read after every replay */
    long long launches;             /* kernels per replay */
    unsigned sites;                 /* peer-memory exchanges per replay */
    bool stale_pads;                /* the cycle pushes level-0 halos: level 0's edge pads are stale after a replay (finalize_pads) */
};

struct sarasmg_solver_s {
    sarasmg_desc d;
    bool planar;
    int D;
    std::vector<Level> lev;
    int vLevel;
    int zeroBC;
    int maxCount;
    cudaStream_t st;
    double *d_wall[3];
    double *d_part;             /* reduction scratch */
    size_t part_cap;
    double *d_scal;             /* small device scalars */
    double *h_scal;             /* pinned host mirror */
    int *d_iters;               /* coarse-solve iteration count */
    int *h_iters;
    double *d_in_p, *d_in_rhs;  /* device-resident inputs of sarasmg_solve_resident */
    /* inter-GPU halo */
    ncclComm_t comm;
    bool has_comm;
    double *xs[2], *xr[2], *ys[2], *yr[2];
    double *d_gather;           /* n_proc * 4 doubles */
    /* peer-memory halo exchange (CUDA IPC over NVLink): receive arena of this rank and the mapped arenas of its face neighbours */
    bool ipc_ok;
    double *ipc_arena;
    size_t ipc_slot, ipc_xc, ipc_yc;        /* doubles per parity slot; level-0 x / y face counts (fix the layout of a slot) */
    double *ipc_peer[4];                    /* mapped arena of the left / right / front / back neighbour (nullptr: none) */
    void *ipc_opened[8];
    int ipc_nopened;
    double *slab;                           /* every level array of this rank: ONE allocation, so one IPC handle maps all of it */
    size_t slab_elems;
    double *slab_peer[4];                   /* the face neighbours' slabs (peer-memory halo push of the Jacobi sweeps) */
    bool push_ok;                           /* the neighbours' slabs are mapped: Jacobi sweeps may store their boundary planes straight into them */
    bool push_on;                           /* inside a V-cycle of mgSolve: intermediate edge pads may stay stale (see op_smooth) */
    bool pads_stale;                        /* level 0's edge pads are behind its core: finalize_pads before the array leaves the solver */
    bool level_zero[32];                    /* pressureData(level) was just set to 0 (p:258) and nothing has written it since */
    unsigned push_add[32];                  /* per level: exchange number (relative to *d_ipc_base) of the push the pads are waiting for; 0 = none */
    unsigned ipc_epoch;                     /* number of the last exchange (dense, the same on every rank) */
    unsigned *d_ipc_base;                   /* device word the exchange kernels add their position to (see IpcPlan) */
    unsigned ipc_base_h;                    /* its current value */
    /* CUDA-graph replay of the V-cycle */
    bool use_graph, graph_multi, capturing;
    unsigned cap_sites;
    std::vector<ProfEntry> *cap_prof;
    std::vector<GraphSlot *> graphs;
    GraphSlot *last_graph;                  /* replayed since the last profile fold */
    long long ipc_cta_max;                  /* faces up to this many doubles are exchanged by the single-CTA kernel */
    /* agglomerated coarsest-level solve (k_coarse_solve_blocks): all ranks' coarsest arrays, block table, metric copies */
    double *agg_P, *agg_R;      /* n_proc * lev[D].elems doubles each */
    CoarseBlock *agg_tab;
    double *agg_met;            /* n_proc * 2 * ((Fx+2) + (Fy+2)) doubles: xixx, xix2, etyy, ety2 of every rank, coarsest level */
    bool agg_ok;
    bool coarse_global;         /* SARASMG_COARSE_GLOBAL: keep the agglomerated coarsest solve in global memory (k_coarse_solve_blocks) */
    bool coarse_legacy;         /* SARASMG_COARSE_LEGACY: host-driven loop, one halo exchange + all-gather per iteration */
    /* instrumentation */
    long long launches;
    std::vector<ProfEntry> pending;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[SARASMG_K_COUNT][32];
    long long prof_n[SARASMG_K_COUNT][32];
    double t_comp, t_tran;
    int bucket;                 /* 0 none, 1 smooth-compute, 2 smooth-transfer */
    cudaEvent_t ev_solve0, ev_solve1;
    double last_solve_ms;
    int last_coarse_iters;
    int mean_override;          /* test hook: use mean_value instead of the device-computed RHS mean */
    double mean_value;
    double last_mean;           /* globalAvg subtracted by the last solve */
    int prof_max_level;         /* events only for operators on levels <= this */
    bool coarsen_legacy;        /* SARASMG_COARSEN_LEGACY: thread-per-coarse-point restriction kernel */
    bool prolong_legacy;        /* SARASMG_PROLONG_LEGACY: thread-per-fine-point prolongation kernel */
    double fuse_min_points;     /* fused residual + restriction on levels with at least this many core points */
    bool no_zero_skip;          /* SARASMG_NO_ZEROSKIP: run the imposeBC of an all-zero coarse-level pressure array too */
    bool no_fuse;               /* SARASMG_NO_FUSE: computeResidual and coarsen as two passes on every level */
    bool no_pow2;               /* SARASMG_NO_POW2: keep the general exact-division path even when the divisors are powers of two (tests) */
    bool gs_legacy;             /* SARASMG_GS_LEGACY: one launch per hyperplane instead of the skewed single-launch sweep */
    cudaEvent_t user_ev[8];
    std::vector<std::pair<const void *, size_t>> pinned;   /* caller ranges this handle page-locked */
    /* device-resident projection step: velocity components, pressure and the first-derivative metrics */
    bool proj_ok;
    ProjView pv;
    double *d_vel[3];           /* Vx, Vy, Vz in their compact Blitz layouts */
    size_t vel_elems[3];
    double *d_presP;            /* P in the pitched level-0 layout */
    double *d_pmet[6];          /* xi_xStaggr, et_yStaggr, zt_zStaggr, xi_xColloc, et_yColloc, zt_zColloc */
    /* implicit velocity solves */
    bool impl_ok;
    sarasmg_impl_desc impl;
    double *d_impl_met[6];      /* xix2Colloc, ety2Colloc, ztz2Colloc, xix2Staggr, ety2Staggr, ztz2Staggr */
    double *d_nse[3];           /* nseRHS components */
    double *d_guess[3];         /* guessedVelocity components (ping-pong partners of d_vel) */
    double *d_psync[2];         /* send / receive buffer of Pp.syncData() between ranks */
    ProjSyncPlan psync;
    int edge_ranks[4];          /* parallel.cc:124-129: always the periodic wrap, never nulled by checkPeriodic */
};

typedef sarasmg_solver_s S;

/* ------------------------------------------------------------------------------------------------ */
/* instrumentation helpers                                                                           */
/* ------------------------------------------------------------------------------------------------ */
/* the stream must be idle: reads the event pairs the last graph replay recorded */
static int graph_prof_fold(S *s) {
    if (!s->last_graph) return 0;
    for (auto &e : s->last_graph->prof) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e.a, e.b));
        s->prof_ms[e.kind][e.level] += ms;
        s->prof_n[e.kind][e.level] += 1;
        if (e.bucket == 1) s->t_comp += ms * 1e-3;
        if (e.bucket == 2) s->t_tran += ms * 1e-3;
    }
    s->last_graph = nullptr;
    return 0;
}

static int prof_fold(S *s) {
    if (s->last_graph) { CU(cudaStreamSynchronize(s->st)); TRY(graph_prof_fold(s)); }
    if (s->pending.empty()) return 0;
    CU(cudaStreamSynchronize(s->st));
    for (auto &e : s->pending) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e.a, e.b));
        s->prof_ms[e.kind][e.level] += ms;
        s->prof_n[e.kind][e.level] += 1;
        if (e.bucket == 1) s->t_comp += ms * 1e-3;
        if (e.bucket == 2) s->t_tran += ms * 1e-3;
        s->ev_pool.push_back(e.a);
        s->ev_pool.push_back(e.b);
    }
    s->pending.clear();
    return 0;
}

struct ProfScope {
    S *s;
    bool on;
    ProfEntry e;
    ProfScope(S *s_, int kind, int level) : s(s_) {
        on = (s->d.flags & (SARASMG_FLAG_PROFILE | SARASMG_FLAG_TIME_RUN)) != 0;
        if (on && !(s->d.flags & SARASMG_FLAG_TIME_RUN) && level > s->prof_max_level) on = false;
        if (!on) return;
        e.kind = kind; e.level = level; e.bucket = s->bucket;
        if (s->capturing) {         /* inside a graph: external event-record nodes, owned by the graph slot */
            cudaEventCreate(&e.a); cudaEventCreate(&e.b);
            cudaEventRecordWithFlags(e.a, s->st, cudaEventRecordExternal);
            return;
        }
        for (cudaEvent_t *ev : {&e.a, &e.b}) {
            if (!s->ev_pool.empty()) { *ev = s->ev_pool.back(); s->ev_pool.pop_back(); }
            else cudaEventCreate(ev);
        }
        cudaEventRecord(e.a, s->st);
    }
    ~ProfScope() {
        if (!on) return;
        if (s->capturing) {
            cudaEventRecordWithFlags(e.b, s->st, cudaEventRecordExternal);
            s->cap_prof->push_back(e);
            return;
        }
        cudaEventRecord(e.b, s->st);
        s->pending.push_back(e);
        if (s->pending.size() > 8192) prof_fold(s);
    }
};

static inline int kcheck(S *s, const char *what) {
    s->launches += 1;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return set_err("kernel launch %s failed: %s", what, cudaGetErrorString(e));
    return 0;
}

/* x is an exact power of two whose reciprocal is a normal double: dividing by x equals multiplying by 1/x bit for bit */
static inline bool is_pow2(double x) {
    int e = 0;
    if (!(x > 0.0) || !std::isfinite(x)) return false;
    return std::frexp(x, &e) == 0.5 && e > -1000 && e < 1000;
}
/* the six divisors of the Laplacian bracket (d3:100-105): hx*hx, 2hx, hy*hy, 2hy, hz*hz, 2hz (2-D: dxx, dzz passed in) */
static inline bool divisors_pow2(bool planar, double dxx, double hx, double hy, double dzz, double hz) {
    return is_pow2(dxx) && is_pow2(2.0 * hx) && is_pow2(dzz) && is_pow2(2.0 * hz) && (planar || (is_pow2(hy * hy) && is_pow2(2.0 * hy)));
}

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }
static inline unsigned face_blocks(long long cnt) { return std::max(1u, std::min(cdiv(cnt, 256), 4096u)); }

/* ------------------------------------------------------------------------------------------------ */
/* host <-> device copies of a level array (compact padded host layout <-> pitched device layout)    */
/* ------------------------------------------------------------------------------------------------ */
static int copy_h2d(S *s, const Level &L, double *dev, const double *host) {
    CU(cudaMemcpy2DAsync(dev + MG_LEAD, (size_t)L.v.pitch * sizeof(double), host, (size_t)L.v.Fz * sizeof(double),
                         (size_t)L.v.Fz * sizeof(double), (size_t)L.v.Fx * L.v.Fy, cudaMemcpyHostToDevice, s->st));
    return 0;
}
static int copy_d2h(S *s, const Level &L, double *host, const double *dev) {
    CU(cudaMemcpy2DAsync(host, (size_t)L.v.Fz * sizeof(double), dev + MG_LEAD, (size_t)L.v.pitch * sizeof(double),
                         (size_t)L.v.Fz * sizeof(double), (size_t)L.v.Fx * L.v.Fy, cudaMemcpyDeviceToHost, s->st));
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* face operations                                                                                   */
/* ------------------------------------------------------------------------------------------------ */
static int launch_faces(S *s, double *A, const Level &L, const int m[6], bool with_walls) {
    bool any = false;
    for (int q = 0; q < 6; q++) any = any || (m[q] != FACE_NONE);
    if (!any) return 0;
    FaceModes F;
    for (int q = 0; q < 6; q++) F.m[q] = m[q];
    if (s->planar) { F.m[2] = FACE_NONE; F.m[3] = FACE_NONE; }
    F.xwall = with_walls ? s->d_wall[0] : nullptr; F.ywall = with_walls ? s->d_wall[1] : nullptr; F.zwall = with_walls ? s->d_wall[2] : nullptr;
    const LevelView &v = L.v;
    const long long nj = s->planar ? 1 : v.Fy;
    const long long cnt = (long long)v.Fx * nj + (s->planar ? 0 : (long long)v.Fx * v.Fz) + nj * v.Fz;
    if (s->planar) k_faces<2><<<face_blocks(cnt), 256, 0, s->st>>>(A, v, F);
    else k_faces<3><<<face_blocks(cnt), 256, 0, s->st>>>(A, v, F);
    return kcheck(s, "k_faces");
}

static int zero_pads(S *s, int level, double *A) {
    const Level &L = s->lev[level];
    ProfScope ps(s, SARASMG_K_MISC, level);
    const int m[6] = {FACE_ZERO, FACE_ZERO, FACE_ZERO, FACE_ZERO, FACE_ZERO, FACE_ZERO};
    return launch_faces(s, A, L, m, false);
}

/* which x / y faces exchange with another rank, and which wrap onto this rank itself */
struct PadPlan { bool rl, rr, rf, rb; int self[6]; bool any_remote, any_self; };
static PadPlan pad_plan(const S *s) {
    PadPlan p;
    const int me = s->d.rank;
    const int *fr = s->d.face_ranks;
    for (int q = 0; q < 6; q++) p.self[q] = FACE_NONE;
    p.rl = (fr[0] != SARASMG_PROC_NULL && fr[0] != me);
    p.rr = (fr[1] != SARASMG_PROC_NULL && fr[1] != me);
    p.rf = !s->planar && (fr[2] != SARASMG_PROC_NULL && fr[2] != me);
    p.rb = !s->planar && (fr[3] != SARASMG_PROC_NULL && fr[3] != me);
    if (fr[0] == me) p.self[0] = FACE_PERIODIC;           /* my i=-1 <- (self) i=xEnd-1 */
    if (fr[1] == me) p.self[1] = FACE_PERIODIC;           /* my i=xEnd+1 <- (self) i=1 */
    if (!s->planar && fr[2] == me) p.self[2] = FACE_PERIODIC;
    if (!s->planar && fr[3] == me) p.self[3] = FACE_PERIODIC;
    p.any_remote = p.rl || p.rr || p.rf || p.rb;
    p.any_self = (p.self[0] | p.self[1] | p.self[2] | p.self[3]) != 0;
    return p;
}

/* ------------------------------------------------------------------------------------------------ */
/* updatePads (d3:671-686, d2:467-478).  x faces first, then y faces (which carry the fresh x pads):  */
/* the program order of the reference's blocking self-sends, made deterministic for real neighbours.  */
/* Neighbour == own rank (periodic, one rank along that axis): local copies (k_faces).                 */
/* Neighbour == other rank: pack -> grouped ncclSend/ncclRecv -> unpack on the solver stream.          */
/* ------------------------------------------------------------------------------------------------ */
/* bc != nullptr: the caller is imposeBC and wants these wall modes applied after the exchange; *bc_done tells it whether
 * the exchange kernel already did (single-CTA path on small levels) */
static int update_pads(S *s, int level, double *A, const int *bc = nullptr, bool *bc_done = nullptr) {
    const Level &L = s->lev[level];
    const LevelView &v = L.v;
    const int *fr = s->d.face_ranks;
    const PadPlan pp = pad_plan(s);
    ProfScope ps(s, SARASMG_K_HALO, level);
    if (!pp.any_remote) return launch_faces(s, A, L, pp.self, false);      /* x then y is the resolution order of k_faces */
    if (!s->has_comm) return set_err("update_pads: remote neighbour but no NCCL communicator");

    const int nj = s->planar ? 1 : v.Fy;
    const long long cx = (long long)nj * v.Fz, cy = (long long)v.Fx * v.Fz;
    if (s->ipc_ok) {
        /* peer-memory path: no library call (kernels and protocol in mg_kernels.cuh).  The exchange number is
         * *d_ipc_base + add: inside a graph capture `add` is the position of the exchange in the graph */
        IpcPlan P;
        const size_t xc = s->ipc_xc, yc = s->ipc_yc;
        unsigned *myflags = (unsigned *)(s->ipc_arena + 2 * s->ipc_slot);
        const bool rem[4] = {pp.rl, pp.rr, pp.rf, pp.rb};
        const size_t send_off[4] = {xc, 0, 2 * xc + yc, 2 * xc}, recv_off[4] = {0, xc, 2 * xc, 2 * xc + yc};
        const int send_side[4] = {1, 0, 3, 2};      /* my left plane lands in the neighbour's RIGHT-side buffer, ... */
        for (int f = 0; f < 4; f++) {
            P.send[f] = rem[f] ? s->ipc_peer[f] + send_off[f] : nullptr;
            P.sflag[f] = rem[f] ? (unsigned *)(s->ipc_peer[f] + 2 * s->ipc_slot) + send_side[f] * 2 : nullptr;
            P.recv[f] = rem[f] ? s->ipc_arena + recv_off[f] : nullptr;
            P.rflag[f] = rem[f] ? myflags + f * 2 : nullptr;
        }
        P.slot_stride = (long long)s->ipc_slot;
        P.base = s->d_ipc_base;
        P.add = s->capturing ? ++s->cap_sites : (++s->ipc_epoch - s->ipc_base_h);
        P.counter = myflags + 16;
        P.err = s->d_ipc_base + 1;
        const bool xrem = pp.rl || pp.rr, yrem = pp.rf || pp.rb;
        if (std::max(cx, s->planar ? 0LL : cy) <= s->ipc_cta_max) {
            /* small level: the whole exchange, and the wall BCs of imposeBC, in one single-CTA launch */
            FaceModes fx, fy, fb;
            for (int q = 0; q < 6; q++) { fx.m[q] = FACE_NONE; fy.m[q] = FACE_NONE; fb.m[q] = bc ? bc[q] : FACE_NONE; }
            fx.m[0] = pp.self[0]; fx.m[1] = pp.self[1]; fy.m[2] = pp.self[2]; fy.m[3] = pp.self[3];
            if (s->planar) { fb.m[2] = FACE_NONE; fb.m[3] = FACE_NONE; }
            fx.xwall = fx.ywall = fx.zwall = fy.xwall = fy.ywall = fy.zwall = nullptr;
            fb.xwall = s->d_wall[0]; fb.ywall = s->d_wall[1]; fb.zwall = s->d_wall[2];
            if (s->planar) k_ipc_exchange_cta<2><<<1, 1024, 0, s->st>>>(A, v, P, fx, fy, fb);
            else k_ipc_exchange_cta<3><<<1, 1024, 0, s->st>>>(A, v, P, fx, fy, fb);
            TRY(kcheck(s, "k_ipc_exchange_cta"));
            if (bc_done) *bc_done = true;
            return 0;
        }
        if (pp.self[0] | pp.self[1]) {
            const int m[6] = {pp.self[0], pp.self[1], FACE_NONE, FACE_NONE, FACE_NONE, FACE_NONE};
            TRY(launch_faces(s, A, L, m, false));
        }
        if (xrem) {
            k_ipc_send_x<<<std::min(face_blocks(cx), 296u), 256, 0, s->st>>>(A, v, s->planar, P);
            TRY(kcheck(s, "k_ipc_send_x"));
        }
        k_ipc_x_to_y<<<std::min(face_blocks(cx + cy), 296u), 256, 0, s->st>>>(A, v, s->planar, P);
        TRY(kcheck(s, "k_ipc_x_to_y"));
        if (pp.self[2] | pp.self[3]) {
            const int m[6] = {FACE_NONE, FACE_NONE, pp.self[2], pp.self[3], FACE_NONE, FACE_NONE};
            TRY(launch_faces(s, A, L, m, false));
        }
        if (yrem) {
            k_ipc_recv_y<<<std::min(face_blocks(cy), 296u), 256, 0, s->st>>>(A, v, P);
            TRY(kcheck(s, "k_ipc_recv_y"));
        }
        return 0;
    }
    /* ---- x phase ---- */
    if (pp.self[0] | pp.self[1]) {
        const int m[6] = {pp.self[0], pp.self[1], FACE_NONE, FACE_NONE, FACE_NONE, FACE_NONE};
        TRY(launch_faces(s, A, L, m, false));
    }
    const bool xrem = pp.rl || pp.rr, yrem = pp.rf || pp.rb;
    if (xrem) {
        k_pack_x2<<<face_blocks(cx), 256, 0, s->st>>>(A, v, s->planar, pp.rl ? s->xs[0] : nullptr, pp.rr ? s->xs[1] : nullptr);
        TRY(kcheck(s, "k_pack_x2"));
        NC(ncclGroupStart());
        /* order: (send left, recv right), (send right, recv left): keeps pairs matched when both neighbours are one peer */
        if (pp.rl) NC(ncclSend(s->xs[0], (size_t)cx, ncclDouble, fr[0], s->comm, s->st));
        if (pp.rr) NC(ncclRecv(s->xr[1], (size_t)cx, ncclDouble, fr[1], s->comm, s->st));
        if (pp.rr) NC(ncclSend(s->xs[1], (size_t)cx, ncclDouble, fr[1], s->comm, s->st));
        if (pp.rl) NC(ncclRecv(s->xr[0], (size_t)cx, ncclDouble, fr[0], s->comm, s->st));
        NC(ncclGroupEnd());
    }
    /* ---- unpack x, pack y (one launch) ---- */
    if (xrem || yrem) {
        k_unpack_x_pack_y<<<face_blocks(cx + cy), 256, 0, s->st>>>(A, v, s->planar, pp.rl ? s->xr[0] : nullptr, pp.rr ? s->xr[1] : nullptr,
                                                                   pp.rf ? s->ys[0] : nullptr, pp.rb ? s->ys[1] : nullptr);
        TRY(kcheck(s, "k_unpack_x_pack_y"));
    }
    /* ---- y phase ---- */
    if (pp.self[2] | pp.self[3]) {
        const int m[6] = {FACE_NONE, FACE_NONE, pp.self[2], pp.self[3], FACE_NONE, FACE_NONE};
        TRY(launch_faces(s, A, L, m, false));
    }
    if (yrem) {
        NC(ncclGroupStart());
        if (pp.rf) NC(ncclSend(s->ys[0], (size_t)cy, ncclDouble, fr[2], s->comm, s->st));
        if (pp.rb) NC(ncclRecv(s->yr[1], (size_t)cy, ncclDouble, fr[3], s->comm, s->st));
        if (pp.rb) NC(ncclSend(s->ys[1], (size_t)cy, ncclDouble, fr[3], s->comm, s->st));
        if (pp.rf) NC(ncclRecv(s->yr[0], (size_t)cy, ncclDouble, fr[2], s->comm, s->st));
        NC(ncclGroupEnd());
        k_unpack_y2<<<face_blocks(cy), 256, 0, s->st>>>(A, v, pp.rf ? s->yr[0] : nullptr, pp.rb ? s->yr[1] : nullptr);
        TRY(kcheck(s, "k_unpack_y2"));
    }
    return 0;
}

/* wall-BC modes of this rank (imposeBC d3:578-667, d2:406-463) */
static void bc_modes(const S *s, int m[6]) {
    const sarasmg_desc &d = s->d;
    const int wall = d.test_poisson ? (s->zeroBC ? FACE_NEG_MIRROR : FACE_WALL) : FACE_MIRROR;
    for (int q = 0; q < 6; q++) m[q] = FACE_NONE;
    if (!d.x_per) {
        if (d.x_rank == 0) m[0] = wall;
        if (d.x_rank == d.npx - 1) m[1] = wall;
    }
    if (!s->planar && !d.y_per) {
        if (d.y_rank == 0) m[2] = wall;
        if (d.y_rank == d.npy - 1) m[3] = wall;
    }
    if (d.z_per) { m[4] = FACE_PERIODIC; m[5] = FACE_PERIODIC; }
    else { m[4] = wall; m[5] = wall; }
}

/* Peer-memory halo PUSH (mg_kernels.cuh, HaloPush).  Inside the V-cycles of mgSolve the Jacobi sweep of a rank with
 * neighbours stores its boundary planes straight into the neighbours' pads, and the imposeBC after it (d3:575-668) only
 * waits for their flags before it applies the wall / periodic-z / self-periodic face operations: no pack, exchange or unpack
 * launches.  A push fills the FACE pads only.  The four edge columns (x pad and y pad at once), which updatePads' x-then-y
 * order fills with the diagonal neighbour's values, keep older data: no 7-point operator reads them; the one reader of
 * edge pads is the 27-point restriction of tmpDataArray, whose exchange stays the full one; and the pads mgSolve hands
 * back (p:151) are rebuilt by one full imposeBC at the end of the solve (finalize_pads) -- imposeBC is a function of the
 * core values only, so that reproduces the reference's last imposeBC exactly. */
static bool push_level_ok(const S *s, int level) {
    if (!s->push_on || !s->push_ok || s->d.gs_smooth || s->planar || level >= 32) return false;
    if (s->d.test_poisson && !s->zeroBC) return false;      /* Dirichlet ghost values: the two-launch wall path stays */
    const int me = s->d.rank;
    const int *fr = s->d.face_ranks;
    for (int f = 0; f < 4; f++) if (fr[f] != SARASMG_PROC_NULL && fr[f] != me) return true;
    return false;
}
static unsigned next_exchange_number(S *s) { return s->capturing ? ++s->cap_sites : (++s->ipc_epoch - s->ipc_base_h); }
static unsigned *my_ipc_flags(const S *s) { return (unsigned *)(s->ipc_arena + 2 * s->ipc_slot); }
enum { IPC_PUSH_FLAG0 = 32, IPC_PUSH_COUNTER = 40 };        /* words of the arena's flag block used by the push (0..17: exchange kernels) */
static HaloPush make_halo_push(S *s, const Level &L, int which) {
    HaloPush H;
    const int me = s->d.rank;
    const int *fr = s->d.face_ranks;
    const int opposite[4] = {1, 0, 3, 2};           /* my left plane is the neighbour's RIGHT pad, ... */
    for (int f = 0; f < 4; f++) {
        const bool rem = fr[f] != SARASMG_PROC_NULL && fr[f] != me;
        H.dst[f] = rem ? s->slab_peer[f] + L.off[which] : nullptr;
        H.sflag[f] = rem ? (unsigned *)(s->ipc_peer[f] + 2 * s->ipc_slot) + IPC_PUSH_FLAG0 + opposite[f] : nullptr;
    }
    H.base = s->d_ipc_base;
    H.add = next_exchange_number(s);
    H.counter = my_ipc_flags(s) + IPC_PUSH_COUNTER;
    return H;
}
/* imposeBC after a pushing sweep: one face launch that first waits for the neighbours' pushes */
static int impose_bc_pushed(S *s, int level) {
    const Level &L = s->lev[level];
    int m[6];
    bc_modes(s, m);
    const PadPlan pp = pad_plan(s);
    FaceModes F;
    for (int q = 0; q < 6; q++) F.m[q] = (m[q] != FACE_NONE) ? m[q] : pp.self[q];
    F.xwall = F.ywall = F.zwall = nullptr;
    HaloWait W;
    const bool rem[4] = {pp.rl, pp.rr, pp.rf, pp.rb};
    for (int f = 0; f < 4; f++) W.flag[f] = rem[f] ? my_ipc_flags(s) + IPC_PUSH_FLAG0 + f : nullptr;
    W.base = s->d_ipc_base;
    W.add = s->push_add[level];
    W.err = s->d_ipc_base + 1;
    s->push_add[level] = 0;
    ProfScope ps(s, SARASMG_K_BC, level);
    const LevelView &v = L.v;
    const long long cnt = (long long)v.Fx * v.Fy + (long long)v.Fx * v.Fz + (long long)v.Fy * v.Fz;
    k_faces_wait<3><<<face_blocks(cnt), 256, 0, s->st>>>(L.buf[SARASMG_PRESSURE], v, F, W);
    return kcheck(s, "k_faces_wait");
}

static int impose_bc(S *s, int level) {
    if (level < 32 && s->push_add[level]) return impose_bc_pushed(s, level);
    const Level &L = s->lev[level];
    double *P = L.buf[SARASMG_PRESSURE];
    int m[6];
    bc_modes(s, m);
    bool has_wall = false;
    for (int q = 0; q < 6; q++) has_wall = has_wall || (m[q] == FACE_WALL);
    if (has_wall && level != 0)
        return set_err("imposeBC: non-homogeneous Dirichlet BC requested on level %d (wall arrays exist on level 0 only)", level);
    const PadPlan pp = pad_plan(s);
    if (!pp.any_remote && !has_wall) {
        /* one launch: without Dirichlet ghost values every operation is an index map (mirror / periodic copy, maybe negated),
         * the maps of different axes commute, and the per-axis source resolution of k_faces gives exactly what the reference's
         * x-pads, y-pads, x-wall, y-wall, z sequence leaves in every pad */
        ProfScope ps(s, SARASMG_K_BC, level);
        int all[6];
        for (int q = 0; q < 6; q++) all[q] = (m[q] != FACE_NONE) ? m[q] : pp.self[q];     /* a wall BC is applied after, and over, the pad copy */
        return launch_faces(s, P, L, all, false);
    }
    bool bc_done = false;
    TRY(update_pads(s, level, P, m, &bc_done));
    if (bc_done) return 0;
    ProfScope ps(s, SARASMG_K_BC, level);
    return launch_faces(s, P, L, m, true);
}

/* ------------------------------------------------------------------------------------------------ */
/* TMA descriptors of a level array (mg_tma.cuh): 3-D fp64 tensor [pitch, Fy, Fx], dense box, zero OOB fill */
/* ------------------------------------------------------------------------------------------------ */
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}
static int make_level_tmaps(S *s, Level &L) {
    L.tma = false;
    const LevelView &v = L.v;
    if (s->planar || getenv("SARASMG_NO_TMA")) return 0;
    if (const char *e = getenv("SARASMG_TMA_MIN")) {            /* tests: force the TMA kernels onto small levels */
        const int m = atoi(e);
        if (v.nx < m || v.ny < m || v.nz < m) return 0;
    } else if (v.nz < 257 || v.ny < 65 || v.nx < 65) return 0;  /* below this the tile ring does not pay; register-marching kernels */
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return set_err("cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)v.pitch, (cuuint64_t)v.Fy, (cuuint64_t)v.Fx};
    const cuuint64_t strides[2] = {(cuuint64_t)v.pitch * sizeof(double), (cuuint64_t)v.plane * sizeof(double)};
    const cuuint32_t boxP[3] = {MG_TMA_TZ + 4, MG_TMA_TY + 2, 1}, boxR[3] = {MG_TMA_TZ, MG_TMA_TY, 1}, es[3] = {1, 1, 1};
    for (int a = 0; a < 3; a++) {
        CUresult r1 = enc(&L.tmapP[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, L.buf[a], dims, strides, boxP, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = enc(&L.tmapR[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, L.buf[a], dims, strides, boxR, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return set_err("cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
        const cuuint32_t boxP16[3] = {MG_FZ + 4, MG_FY + 2, 1}, boxR16[3] = {MG_FZ, MG_FY, 1};
        r1 = enc(&L.tmapP16[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, L.buf[a], dims, strides, boxP16, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        r2 = enc(&L.tmapR16[a], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, L.buf[a], dims, strides, boxR16, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return set_err("cuTensorMapEncodeTiled (16-row tiles) failed (%d, %d)", (int)r1, (int)r2);
    }
    L.tma = true;
    return 0;
}
struct TmaGeom { dim3 grid, block; int ichunk; };
static TmaGeom tma_geom(const LevelView &v) {
    TmaGeom g;
    g.block = dim3(MG_TMA_TZ / 2, MG_TMA_TY / 2, 1);      /* a thread owns a z pair of two adjacent rows */
    const unsigned gx = cdiv(v.nz, MG_TMA_TZ), gy = cdiv(v.ny, MG_TMA_TY);
    /* x chunks: enough CTAs for ~8 waves of 2 CTAs on 148 SMs, as few as possible otherwise (each chunk re-reads 2 planes) */
    int nch = (int)cdiv(148LL * 2 * 8, (long long)gx * gy);
    nch = std::max(1, std::min(nch, v.nx / 8));
    nch = std::max(nch, (v.nx + MG_TMA_MAXCH - 1) / MG_TMA_MAXCH);     /* the kernel's x-coefficient table holds MG_TMA_MAXCH planes */
    g.ichunk = (v.nx + nch - 1) / nch;
    g.grid = dim3(gx, gy, cdiv(v.nx, g.ichunk));
    return g;
}

/* ------------------------------------------------------------------------------------------------ */
/* marching-kernel launch geometry                                                                   */
/* ------------------------------------------------------------------------------------------------ */
struct MarchGeom { dim3 grid, block; int ichunk; };

static MarchGeom march_geom(const S *s, const LevelView &v) {
    MarchGeom g;
    const int pairs = (v.nz + 1) / 2;
    int bx = 32, by = 8;
    if (s->planar) { bx = 128; by = 1; }
    else {
        while (bx > 4 && bx / 2 >= pairs) { bx /= 2; by *= 2; }
    }
    g.block = dim3(bx, by, 1);
    const unsigned gx = cdiv(pairs, bx), gy = s->planar ? 1 : cdiv(v.ny, by);
    /* pick the x-chunk so that the grid has >= ~8 CTAs per SM (148 SMs) when the level is big enough */
    int ichunk = 64;
    while (ichunk > 4 && (long long)gx * gy * cdiv(v.nx, ichunk) < 148LL * 8) ichunk /= 2;
    g.ichunk = ichunk;
    g.grid = dim3(gx, gy, cdiv(v.nx, ichunk));
    return g;
}

/* ------------------------------------------------------------------------------------------------ */
/* operators                                                                                         */
/* ------------------------------------------------------------------------------------------------ */
/* builds the task table of the skewed Gauss-Seidel kernel for one level (gs_skew.cuh) */
static int build_gs_plan(S *s, Level &L) {
    const LevelView &v = L.v;
    const int extA = s->planar ? 1 : v.nx, extB = s->planar ? v.nx : v.ny;
    GsPlan &G = L.gs;
    int ti = 1;
    if (!s->planar) {
        int want = 6;                /* 6 compute warps + halo warp + publisher warp = 256 threads: 128 registers at 2 CTAs per SM */
        if (const char *e = getenv("SARASMG_GS_TI")) want = std::max(1, std::min(6, atoi(e)));
        const int nt = (extA + want - 1) / want;
        ti = (extA + nt - 1) / nt;
    }
    G.TI = ti;
    G.NI = (extA + ti - 1) / ti;
    G.NJ = (extB + GS_TJ - 1) / GS_TJ;
    G.NQ = (v.nz + 3) / 4;
    int zq = 128;
    if (const char *e = getenv("SARASMG_GS_ZQ")) zq = std::max(1, atoi(e));
    G.ZQ = std::min(zq, G.NQ);
    G.NZ = std::max(1, G.NQ / G.ZQ);
    G.ntasks = G.NI * G.NJ * G.NZ;
    G.tabn = 4 * (G.NQ - (G.NZ - 1) * G.ZQ);
    G.z_uniform = L.z_uniform ? 1 : 0;
    G.slack = 0;
    if (const char *e = getenv("SARASMG_GS_SLACK")) G.slack = std::max(0, atoi(e));     /* tuning knob, see GsPlan::slack */
    /* ticket order: ideal start time of a task in the systolic schedule; predecessors (I-1,J,Z), (I,J-1,Z), (I,J,Z-1) sort first */
    struct T { long long key; int I, J, Z; };
    std::vector<T> ts;
    ts.reserve(G.ntasks);
    const long long zspan = (long long)2 * G.ZQ + ti + GS_TJ;
    for (int Z = 0; Z < G.NZ; Z++)
        for (int J = 0; J < G.NJ; J++)
            for (int I = 0; I < G.NI; I++) ts.push_back({(long long)ti * I + (long long)GS_TJ * J + zspan * Z, I, J, Z});
    std::stable_sort(ts.begin(), ts.end(), [](const T &a, const T &b) { return a.key < b.key; });
    std::vector<int4> ht(G.ntasks);
    for (int q = 0; q < G.ntasks; q++) ht[q] = make_int4(ts[q].I, ts[q].J, ts[q].Z, 0);
    CU(cudaMalloc(&L.gs_tasks, sizeof(int4) * G.ntasks));
    CU(cudaMemcpyAsync(L.gs_tasks, ht.data(), sizeof(int4) * G.ntasks, cudaMemcpyHostToDevice, s->st));
    CU(cudaMalloc(&L.gs_sync, sizeof(int) * (G.ntasks + 1)));
    CU(cudaStreamSynchronize(s->st));
    G.tasks = L.gs_tasks;
    G.sync = L.gs_sync;
    L.gs_smem = sizeof(double) * ((size_t)2 * GS_DP * (ti + 2) * GS_ROW * 4 + (size_t)GS_DP * ti * 32 * 4 + (size_t)2 * G.tabn);
    return 0;
}

/* geometry of the warp-strip Gauss-Seidel kernel for one level (gs_strip.cuh): quads per slab, prefetch distance and warps
 * per CTA are chosen so that every tile is resident at once (one CTA per SM: a tile holds its SM for the whole sweep, a
 * second wave would double the sweep time) with as many compute warps as possible */
static int build_gs_strip_plan(S *s, Level &L) {
    L.gss_ok = false; L.gss_tasks = nullptr; L.gss_sync = nullptr; L.gss_rden = nullptr; L.gss_xpair = nullptr; L.gss_smem = 0;
    const LevelView &v = L.v;
    if (s->planar) return 0;
    /* OPT-IN (SARASMG_GS_STRIP=1).  Bit-exact, but measured slower than k_gs_skew on B200 (1025^3 sweep: 15-21 ms against
     * 11 ms, profiles/r02_gs_strip_*.txt): both kernels are bound by warps-per-SM / instruction latency, not by HBM, and the
     * window rings of this one leave room for only 11-13 warps per SM.  Kept for the analysis in DESIGN.md and as the
     * starting point of the next attempt. */
    {
        const char *e = getenv("SARASMG_GS_STRIP");
        if (!e || atoi(e) == 0) return 0;
    }
    int min_n = 200;                                    /* below this the start-up skew of the strip pipeline outweighs its gain */
    if (const char *e = getenv("SARASMG_GS_STRIP_MIN")) min_n = atoi(e);
    if (v.nx < min_n || v.ny < 32 || v.nz < 16) return 0;
    int dev = 0, nsm = 148, smem_max = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int budget = smem_max - 1024;                 /* static control block + alignment */
    const int NQ = (v.nz + 3) / 4, NJ = (v.ny + 31) / 32;
    const int f_nq = getenv("SARASMG_GSS_NQ") ? atoi(getenv("SARASMG_GSS_NQ")) : 0;
    const int f_dp = getenv("SARASMG_GSS_DP") ? atoi(getenv("SARASMG_GSS_DP")) : 0;
    const int f_wc = getenv("SARASMG_GSS_WC") ? atoi(getenv("SARASMG_GSS_WC")) : 0;
    long long best = -1;
    GsStripPlan B;
    memset(&B, 0, sizeof(B));
    for (int DP = 3; DP >= 2; DP--) {
        if (f_dp && DP != f_dp) continue;
        for (int nq = 1; nq <= 16; nq++) {
            if (f_nq && nq != f_nq) continue;
            const int bytes = gss_warp_bytes(nq, DP);
            int wcmax = std::min(GSS_MAXW, budget / bytes);
            if (f_wc) wcmax = std::min(wcmax, f_wc);
            if (wcmax < 1) continue;
            const int NZW = (NQ + nq - 1) / nq;
            const int NG = (NZW + wcmax - 1) / wcmax;
            const int WC = (NZW + NG - 1) / NG;
            if ((long long)NJ * NG > nsm && !f_nq) continue;
            /* more warps hide more latency; among equal counts prefer the deeper prefetch, then the taller slab */
            const long long score = (long long)NJ * NZW * 1000 + DP * 20 + nq;
            if (score > best) {
                best = score;
                B.NJ = NJ; B.NZW = NZW; B.NG = NG; B.nq = nq; B.NQ = NQ; B.WC = WC; B.DP = DP;
                B.RW = 2 * nq + DP + 1; B.RRn = DP + 1; B.ntasks = NJ * NG;
                B.warp_bytes = gss_warp_bytes(nq, DP);
            }
        }
    }
    if (best < 0) return 0;
    B.z_uniform = L.zz_uniform ? 1 : 0;
    B.two_h = 2.0 * v.hxhyhz;
    CU(cudaMalloc(&L.gss_xpair, sizeof(double) * 2 * (size_t)v.nx));
    k_gss_xpair<<<cdiv(v.nx, 256), 256, 0, s->st>>>(v, L.gss_xpair);
    TRY(kcheck(s, "k_gss_xpair"));
    B.xpair = L.gss_xpair;
    B.rden = nullptr;
    if (B.z_uniform) {
        CU(cudaMalloc(&L.gss_rden, sizeof(double) * (size_t)v.nx * v.ny));
        k_gss_rden<<<std::min(cdiv((long long)v.nx * v.ny, 256), 2048u), 256, 0, s->st>>>(v, L.gss_rden);
        TRY(kcheck(s, "k_gss_rden"));
        B.rden = L.gss_rden;
    }
    /* ticket order: ideal start step of a tile (strip J starts 32 steps after J-1, a slab nq steps after the one below) */
    struct T { long long key; int J, G; };
    std::vector<T> ts;
    for (int g = 0; g < B.NG; g++)
        for (int J = 0; J < B.NJ; J++) ts.push_back({32LL * J + (long long)B.nq * B.WC * g, J, g});
    std::stable_sort(ts.begin(), ts.end(), [](const T &a, const T &b) { return a.key < b.key; });
    std::vector<int2> ht(B.ntasks);
    for (int q = 0; q < B.ntasks; q++) ht[q] = make_int2(ts[q].J, ts[q].G);
    CU(cudaMalloc(&L.gss_tasks, sizeof(int2) * B.ntasks));
    CU(cudaMemcpyAsync(L.gss_tasks, ht.data(), sizeof(int2) * B.ntasks, cudaMemcpyHostToDevice, s->st));
    CU(cudaMalloc(&L.gss_sync, sizeof(int) * ((size_t)B.NJ * B.NZW + 1)));
    CU(cudaStreamSynchronize(s->st));
    B.tasks = L.gss_tasks; B.sync = L.gss_sync;
    L.gss = B;
    L.gss_smem = (size_t)B.WC * B.warp_bytes;
    L.gss_ok = true;
    if (getenv("SARASMG_VERBOSE") && s->d.rank == 0)
        fprintf(stderr, "saras-mg: strip GS on %dx%dx%d: %d strips x %d slabs of %d quads, %d CTAs x %d warps, DP %d, %zu B smem\n",
                v.nx, v.ny, v.nz, B.NJ, B.NZW, B.nq, B.ntasks, B.WC, B.DP, L.gss_smem);
    return 0;
}

static int gs_sweep(S *s, int level) {
    Level &L = s->lev[level];
    const LevelView &v = L.v;
    double *P = L.buf[SARASMG_PRESSURE];
    const double *R = L.buf[SARASMG_RESIDUAL];
    if (!s->gs_legacy && L.gss_ok) {
        /* large 3-D level: warp-strip pipeline (gs_strip.cuh) */
        const GsStripPlan &G = L.gss;
        CU(cudaMemsetAsync(L.gss_sync, 0, sizeof(int) * ((size_t)G.NJ * G.NZW + 1), s->st));
        if (G.z_uniform) k_gs_strip<true><<<G.ntasks, dim3(32, G.WC + 1, 1), L.gss_smem, s->st>>>(P, R, v, G);
        else k_gs_strip<false><<<G.ntasks, dim3(32, G.WC + 1, 1), L.gss_smem, s->st>>>(P, R, v, G);
        return kcheck(s, "k_gs_strip");
    }
    if (!s->gs_legacy) {
        /* one launch per sweep: skewed pipeline with ticketed tile tasks (gs_skew.cuh) */
        const GsPlan &G = L.gs;
        CU(cudaMemsetAsync(L.gs_sync, 0, sizeof(int) * (G.ntasks + 1), s->st));
        const dim3 block(32, G.TI + 2, 1);
        if (s->planar) k_gs_skew<2><<<G.ntasks, block, L.gs_smem, s->st>>>(P, R, v, G);
        else k_gs_skew<3><<<G.ntasks, block, L.gs_smem, s->st>>>(P, R, v, G);
        return kcheck(s, "k_gs_skew");
    }
    if (!s->planar) {
        const int H = v.nx + v.ny + v.nz - 2;
        for (int h = 0; h < H; ++h) {
            const int ilo = std::max(0, h - (v.ny - 1) - (v.nz - 1));
            const int ihi = std::min(v.nx - 1, h);
            dim3 block(32, 8, 1);
            dim3 grid(cdiv(v.ny, 32), cdiv(ihi - ilo + 1, 8), 1);
            k_gs_plane<3><<<grid, block, 0, s->st>>>(P, R, v, h, ilo);
            TRY(kcheck(s, "k_gs_plane<3>"));
        }
    } else {
        const int H = v.nx + v.nz - 1;
        for (int h = 0; h < H; ++h) {
            const int ilo = std::max(0, h - (v.nz - 1));
            const int ihi = std::min(v.nx - 1, h);
            k_gs_plane<2><<<cdiv(ihi - ilo + 1, 128), 128, 0, s->st>>>(P, R, v, h, ilo);
            TRY(kcheck(s, "k_gs_plane<2>"));
        }
    }
    return 0;
}

/* smooth (d3:114-187, d2:109-147) */
static int op_smooth(S *s, int count) {
    const int l = s->vLevel;
    Level &L = s->lev[l];
    /* pressureData(l) == 0 everywhere, pads included (the memset of p:258 just ran): with homogeneous mirror / periodic faces
     * the first imposeBC copies zeros onto zeros on every rank -- skipped.  (Not under TEST_POISSON, whose odd mirror
     * would write -0.0.) */
    const bool skip_first_bc = l < 32 && s->level_zero[l] && s->zeroBC && !s->d.test_poisson && count > 0 && !s->no_zero_skip;
    if (l < 32) s->level_zero[l] = false;
    for (int n = 0; n < count; ++n) {
        s->bucket = 2;
        if (!(n == 0 && skip_first_bc)) TRY(impose_bc(s, l));
        s->bucket = 1;
        {
            ProfScope ps(s, SARASMG_K_SMOOTH, l);
            if (s->d.gs_smooth) {
                TRY(gs_sweep(s, l));
            } else {
                const MarchGeom g = march_geom(s, L.v);
                double *P = L.buf[SARASMG_PRESSURE], *R = L.buf[SARASMG_RESIDUAL], *T = L.buf[SARASMG_TMP];
                if (push_level_ok(s, l)) {
                    /* rank with neighbours, inside mgSolve: the sweep stores its boundary planes into the neighbours' pads */
                    const HaloPush H = make_halo_push(s, L, SARASMG_TMP);
                    s->push_add[l] = H.add;
                    if (l == 0) s->pads_stale = true;
                    if (L.tma) {
                        const TmaGeom t = tma_geom(L.v);
                        if (L.x_uniform) k_stencil_tma_push<true><<<t.grid, t.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], T, L.v, t.ichunk, H);
                        else k_stencil_tma_push<false><<<t.grid, t.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], T, L.v, t.ichunk, H);
                    } else {
                        if (L.x_uniform) k_stencil_march_push<true><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk, H);
                        else k_stencil_march_push<false><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk, H);
                    }
                } else if (L.tma) {
                    const TmaGeom t = tma_geom(L.v);
                    if (L.x_uniform) k_stencil_tma<0, true><<<t.grid, t.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], T, L.v, t.ichunk, 0, 0.0, 0.0, 0.0, nullptr, 0);
                    else k_stencil_tma<0, false><<<t.grid, t.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], T, L.v, t.ichunk, 0, 0.0, 0.0, 0.0, nullptr, 0);
                } else if (s->planar) {
                    if (L.x_uniform) k_stencil_march<2, 0, true><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
                    else k_stencil_march<2, 0, false><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
                } else {
                    if (L.x_uniform) k_stencil_march<3, 0, true><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
                    else k_stencil_march<3, 0, false><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
                }
                TRY(kcheck(s, "k_stencil_march<jacobi>"));
                std::swap(L.buf[SARASMG_PRESSURE], L.buf[SARASMG_TMP]);      /* d3:168 */
                std::swap(L.off[SARASMG_PRESSURE], L.off[SARASMG_TMP]);
                std::swap(L.tmapP[SARASMG_PRESSURE], L.tmapP[SARASMG_TMP]);
                std::swap(L.tmapR[SARASMG_PRESSURE], L.tmapR[SARASMG_TMP]);
                std::swap(L.tmapP16[SARASMG_PRESSURE], L.tmapP16[SARASMG_TMP]);
                std::swap(L.tmapR16[SARASMG_PRESSURE], L.tmapR16[SARASMG_TMP]);
            }
        }
        s->bucket = 0;
    }
    s->bucket = 2;
    TRY(impose_bc(s, l));
    s->bucket = 0;
    return 0;
}

/* computeResidual (d3:90-111, d2:90-106) */
static int op_compute_residual(S *s) {
    const int l = s->vLevel;
    Level &L = s->lev[l];
    TRY(zero_pads(s, l, L.buf[SARASMG_TMP]));
    {
        ProfScope ps(s, SARASMG_K_RESIDUAL, l);
        const MarchGeom g = march_geom(s, L.v);
        double *P = L.buf[SARASMG_PRESSURE], *R = L.buf[SARASMG_RESIDUAL], *T = L.buf[SARASMG_TMP];
        const bool p2 = divisors_pow2(s->planar, L.v.hx * L.v.hx, L.v.hx, L.v.hy, L.v.hz * L.v.hz, L.v.hz) && !s->no_pow2;
        if (L.tma) {
            const TmaGeom t = tma_geom(L.v);
            if (p2) k_stencil_tma<1, true><<<t.grid, t.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], T, L.v, t.ichunk, 0, 0.0, 0.0, 0.0, nullptr, 0);
            else k_stencil_tma<1, false><<<t.grid, t.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], T, L.v, t.ichunk, 0, 0.0, 0.0, 0.0, nullptr, 0);
        } else if (s->planar) {
            if (p2) k_stencil_march<2, 1, true><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
            else k_stencil_march<2, 1, false><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
        } else {
            if (p2) k_stencil_march<3, 1, true><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
            else k_stencil_march<3, 1, false><<<g.grid, g.block, 0, s->st>>>(P, R, T, L.v, g.ichunk);
        }
        TRY(kcheck(s, "k_stencil_march<residual>"));
    }
    TRY(update_pads(s, l, L.buf[SARASMG_TMP]));
    return 0;
}

/* coarsen (d3:250-298, d2:202-236): vLevel += 1 */
static int op_coarsen(S *s) {
    const int pLevel = s->vLevel;
    if (pLevel >= s->D) return set_err("coarsen: already at the coarsest level");
    s->vLevel += 1;
    Level &Lf = s->lev[pLevel];
    Level &Lc = s->lev[s->vLevel];
    ProfScope ps(s, SARASMG_K_COARSEN, pLevel);
    if (s->planar) {
        k_coarsen<2><<<dim3(cdiv(Lc.v.nz, 128), 1, Lc.v.nx), dim3(128, 1, 1), 0, s->st>>>(Lf.buf[SARASMG_TMP], Lf.v, Lc.buf[SARASMG_RESIDUAL], Lc.v);
    } else if (s->coarsen_legacy) {
        int bx = 64, by = 4;
        while (bx > 4 && bx / 2 >= Lc.v.nz) { bx /= 2; by *= 2; }
        k_coarsen<3><<<dim3(cdiv(Lc.v.nz, bx), cdiv(Lc.v.ny, by), Lc.v.nx), dim3(bx, by, 1), 0, s->st>>>(Lf.buf[SARASMG_TMP], Lf.v, Lc.buf[SARASMG_RESIDUAL], Lc.v);
    } else {
        int bx = 32, by = 8;
        while (bx > 4 && bx / 2 >= Lc.v.nz) { bx /= 2; by *= 2; }
        const unsigned gx = cdiv(Lc.v.nz, bx), gy = cdiv(Lc.v.ny, by);
        int ichunk = 32;
        while (ichunk > 2 && (long long)gx * gy * cdiv(Lc.v.nx, ichunk) < 148LL * 8) ichunk /= 2;
        k_coarsen_march<<<dim3(gx, gy, cdiv(Lc.v.nx, ichunk)), dim3(bx, by, 1), 0, s->st>>>(Lf.buf[SARASMG_TMP], Lf.v, Lc.buf[SARASMG_RESIDUAL], Lc.v, ichunk);
    }
    return kcheck(s, "k_coarsen");
}

/* computeResidual + coarsen of the V-cycle (p:250-254) as one pass on the TMA levels (mg_fused.cuh): interior coarse points
 * straight from pressure and rhs, the coarse faces through a two-plane shell of tmpDataArray with the usual pad handling.
 * tmpDataArray's interior is NOT written: nothing reads it before the next sweep / residual overwrites it. */
static bool fuse_level_ok(const S *s, int level) {
    if (s->planar || s->no_fuse || level >= s->D) return false;
    const Level &Lf = s->lev[level], &Lc = s->lev[level + 1];
    if (!(Lf.tma && Lc.v.nx >= 3 && Lc.v.ny >= 3 && Lc.v.nz >= 3 && Lf.v.nx >= 5 && Lf.v.ny >= 5 && Lf.v.nz >= 5)) return false;
    /* measured (mg_fused.cuh): the fused pass wins on the largest levels only; SARASMG_FUSE_MIN overrides the threshold (tests: 0) */
    return (double)Lf.v.nx * Lf.v.ny * Lf.v.nz >= s->fuse_min_points;
}
static int op_residual_coarsen_fused(S *s) {
    const int l = s->vLevel;
    Level &Lf = s->lev[l];
    Level &Lc = s->lev[l + 1];
    double *P = Lf.buf[SARASMG_PRESSURE], *R = Lf.buf[SARASMG_RESIDUAL], *T = Lf.buf[SARASMG_TMP];
    const LevelView &v = Lf.v, &c = Lc.v;
    const bool p2 = divisors_pow2(false, v.hx * v.hx, v.hx, v.hy, v.hz * v.hz, v.hz) && !s->no_pow2;
    TRY(zero_pads(s, l, T));
    {
        ProfScope ps(s, SARASMG_K_FUSED_RC, l);
        const long long cnt = 4LL * v.ny * v.nz + 4LL * (v.nx - 4) * v.nz + 4LL * (v.nx - 4) * (v.ny - 4);
        if (p2) k_resid_shell<true><<<face_blocks(cnt), 256, 0, s->st>>>(P, R, T, v);
        else k_resid_shell<false><<<face_blocks(cnt), 256, 0, s->st>>>(P, R, T, v);
        TRY(kcheck(s, "k_resid_shell"));
        /* interior: independent of the shell and of every pad */
        int CI = MG_F_MAXCI;
        const unsigned gx = cdiv(c.nz - 2, MG_F_CK), gy = cdiv(c.ny - 2, MG_F_CJ);
        while (CI > 4 && (long long)gx * gy * cdiv(c.nx - 2, CI) < 148LL * 4) CI = (CI + 1) / 2;
        const dim3 grid(gx, gy, cdiv(c.nx - 2, CI)), block(MG_FZ / 2, MG_FY / 2, 1);
        if (p2) k_resid_coarsen_tma<true><<<grid, block, MG_F_SMEM, s->st>>>(Lf.tmapP16[SARASMG_PRESSURE], Lf.tmapR16[SARASMG_RESIDUAL], v, Lc.buf[SARASMG_RESIDUAL], c, CI);
        else k_resid_coarsen_tma<false><<<grid, block, MG_F_SMEM, s->st>>>(Lf.tmapP16[SARASMG_PRESSURE], Lf.tmapR16[SARASMG_RESIDUAL], v, Lc.buf[SARASMG_RESIDUAL], c, CI);
        TRY(kcheck(s, "k_resid_coarsen_tma"));
    }
    TRY(update_pads(s, l, T));
    s->vLevel += 1;
    {
        ProfScope ps(s, SARASMG_K_FUSED_RC, l);
        const long long cnt = 2LL * c.ny * c.nz + 2LL * (c.nx - 2) * c.nz + 2LL * (c.nx - 2) * (c.ny - 2);
        k_coarsen_shell<<<face_blocks(cnt), 256, 0, s->st>>>(T, v, Lc.buf[SARASMG_RESIDUAL], c);
        TRY(kcheck(s, "k_coarsen_shell"));
    }
    return 0;
}

/* prolong (d3:301-365, d2:239-271) fused with pressureData += smoothedPres (p:271): vLevel -= 1 */
static int op_prolong(S *s) {
    const int pLevel = s->vLevel;
    if (pLevel <= 0) return set_err("prolong: already at the finest level");
    s->vLevel -= 1;
    Level &Lc = s->lev[pLevel];
    Level &Lf = s->lev[s->vLevel];
    ProfScope ps(s, SARASMG_K_PROLONG, s->vLevel);
    if (s->prolong_legacy) {
        if (s->planar) {
            k_prolong_add<2><<<dim3(cdiv(Lf.v.nz, 128), 1, Lf.v.nx), dim3(128, 1, 1), 0, s->st>>>(Lc.buf[SARASMG_PRESSURE], Lc.v, Lf.buf[SARASMG_PRESSURE], Lf.v);
        } else {
            int bx = 64, by = 4;
            while (bx > 4 && bx / 2 >= Lf.v.nz) { bx /= 2; by *= 2; }
            k_prolong_add<3><<<dim3(cdiv(Lf.v.nz, bx), cdiv(Lf.v.ny, by), Lf.v.nx), dim3(bx, by, 1), 0, s->st>>>(Lc.buf[SARASMG_PRESSURE], Lc.v, Lf.buf[SARASMG_PRESSURE], Lf.v);
        }
        return kcheck(s, "k_prolong_add");
    }
    {   /* marching kernel: threads over the coarse (k2, j2) cells, x in chunks sized for >= ~8 CTAs per SM */
        int bx = 32, by = 8;
        if (s->planar) { bx = 128; by = 1; }
        else while (bx > 4 && bx / 2 >= Lc.v.nz) { bx /= 2; by *= 2; }
        const unsigned gx = cdiv(Lc.v.nz, bx), gy = s->planar ? 1 : cdiv(Lc.v.ny, by);
        int ichunk = 32;
        while (ichunk > 2 && (long long)gx * gy * cdiv(Lc.v.nx, ichunk) < 148LL * 8) ichunk /= 2;
        const dim3 grid(gx, gy, cdiv(Lc.v.nx, ichunk)), block(bx, by, 1);
        if (s->planar) k_prolong_march<2><<<grid, block, 0, s->st>>>(Lc.buf[SARASMG_PRESSURE], Lc.v, Lf.buf[SARASMG_PRESSURE], Lf.v, ichunk);
        else k_prolong_march<3><<<grid, block, 0, s->st>>>(Lc.buf[SARASMG_PRESSURE], Lc.v, Lf.buf[SARASMG_PRESSURE], Lf.v, ichunk);
    }
    return kcheck(s, "k_prolong_march");
}

static int ensure_part(S *s, size_t n) {
    if (n <= s->part_cap) return 0;
    if (s->d_part) CU(cudaFree(s->d_part));
    s->part_cap = std::max(n, (size_t)1 << 16);
    CU(cudaMalloc(&s->d_part, s->part_cap * sizeof(double)));
    return 0;
}

/* gather `nq` per-rank doubles (already in d_scal[0..nq)) from every rank into h (n_proc*nq), rank order */
static int gather_scalars(S *s, int nq, std::vector<double> &out) {
    const int np = s->d.n_proc;
    out.assign((size_t)np * nq, 0.0);
    if (np == 1 || !s->has_comm) {
        CU(cudaMemcpyAsync(s->h_scal, s->d_scal, nq * sizeof(double), cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        for (int q = 0; q < nq; q++) out[q] = s->h_scal[q];
        return 0;
    }
    NC(ncclAllGather(s->d_scal, s->d_gather, (size_t)nq, ncclDouble, s->comm, s->st));
    CU(cudaMemcpyAsync(s->h_scal, s->d_gather, (size_t)np * nq * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (int q = 0; q < np * nq; q++) out[q] = s->h_scal[q];
    return 0;
}

/* computeError (d3:391-459, d2:274-338); sums over ranks in rank order */
static int op_compute_error(S *s, int norm, double *value) {
    *value = 0.0;
    if (norm != 1 && norm != 2) return 0;           /* resType 0: returns 0 (d3:392,418) */
    Level &L = s->lev[0];
    const MarchGeom g = march_geom(s, L.v);
    const TmaGeom tg = tma_geom(L.v);
    const int nb = L.tma ? tg.grid.x * tg.grid.y * tg.grid.z : g.grid.x * g.grid.y * g.grid.z;
    TRY(ensure_part(s, (size_t)2 * nb + 4096));
    {
        ProfScope ps(s, SARASMG_K_ERROR, 0);
        const Level &Lv = s->lev[s->vLevel];        /* d3:411 uses hx(vLevel) */
        double *P = L.buf[SARASMG_PRESSURE], *R = L.buf[SARASMG_RESIDUAL];
        if (s->planar) {
            const bool p2 = divisors_pow2(true, L.v.hx2, L.v.hx, 0.0, L.v.hz2, L.v.hz) && !s->no_pow2;
            if (p2) k_error_march<2, true><<<g.grid, g.block, 0, s->st>>>(P, R, L.v, g.ichunk, norm, L.v.hx, L.v.hy, L.v.hz, s->d_part, nb);
            else k_error_march<2, false><<<g.grid, g.block, 0, s->st>>>(P, R, L.v, g.ichunk, norm, L.v.hx, L.v.hy, L.v.hz, s->d_part, nb);
        } else if (L.tma) {
            const bool p2 = divisors_pow2(false, Lv.v.hx * Lv.v.hx, Lv.v.hx, Lv.v.hy, Lv.v.hz * Lv.v.hz, Lv.v.hz) && !s->no_pow2;
            if (p2) k_stencil_tma<2, true><<<tg.grid, tg.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], nullptr, L.v, tg.ichunk, norm, Lv.v.hx, Lv.v.hy, Lv.v.hz, s->d_part, nb);
            else k_stencil_tma<2, false><<<tg.grid, tg.block, MG_TMA_SMEM, s->st>>>(L.tmapP[SARASMG_PRESSURE], L.tmapR[SARASMG_RESIDUAL], nullptr, L.v, tg.ichunk, norm, Lv.v.hx, Lv.v.hy, Lv.v.hz, s->d_part, nb);
        } else {
            const bool p2 = divisors_pow2(false, Lv.v.hx * Lv.v.hx, Lv.v.hx, Lv.v.hy, Lv.v.hz * Lv.v.hz, Lv.v.hz) && !s->no_pow2;
            if (p2) k_error_march<3, true><<<g.grid, g.block, 0, s->st>>>(P, R, L.v, g.ichunk, norm, Lv.v.hx, Lv.v.hy, Lv.v.hz, s->d_part, nb);
            else k_error_march<3, false><<<g.grid, g.block, 0, s->st>>>(P, R, L.v, g.ichunk, norm, Lv.v.hx, Lv.v.hy, Lv.v.hz, s->d_part, nb);
        }
        TRY(kcheck(s, "k_error_march"));
        k_reduce_final<<<1, 1024, 0, s->st>>>(s->d_part, nb, 2, (norm == 1) ? 1u : 0u, s->d_scal);
        TRY(kcheck(s, "k_reduce_final"));
        if (norm == 1) {
            /* denominator: max |rhs| over the whole padded array (d3:437) */
            const unsigned nbp = std::min(cdiv((long long)L.v.Fx * L.v.Fy, 4), 2048u);
            k_reduce_padded<1><<<nbp, 256, 0, s->st>>>(L.buf[SARASMG_RESIDUAL], L.v, s->d_part);
            TRY(kcheck(s, "k_reduce_padded"));
            k_reduce_final<<<1, 1024, 0, s->st>>>(s->d_part, (int)nbp, 1, 1u, s->d_scal + 1);
            TRY(kcheck(s, "k_reduce_final"));
        }
    }
    std::vector<double> g2;
    TRY(gather_scalars(s, 2, g2));
    const int np = s->d.n_proc;
    const long long cnt_loc = (long long)L.v.nx * L.v.ny * L.v.nz;
    if (norm == 1) {
        double num = g2[0], den = g2[1];
        for (int r = 1; r < np; r++) { num = std::max(num, g2[2 * r]); den = std::max(den, g2[2 * r + 1]); }
        *value = den ? num / den : num;
    } else {
        double num = 0.0, den = 0.0;
        for (int r = 0; r < np; r++) { num += g2[2 * r]; den += g2[2 * r + 1]; }
        const int cnt = (int)(cnt_loc * np);        /* valCountGlo is an int in the reference (d3:434) */
        *value = den ? std::sqrt(num / cnt) / std::sqrt(den / cnt) : std::sqrt(num / cnt);
    }
    return 0;
}

static void bc_plan(const S *s, BcPlan &B) {
    int m[6];
    bc_modes(s, m);
    const int me = s->d.rank;
    const int *fr = s->d.face_ranks;
    B.pad_x_lo = (fr[0] == me) ? FACE_PERIODIC : FACE_NONE;
    B.pad_x_hi = (fr[1] == me) ? FACE_PERIODIC : FACE_NONE;
    B.pad_y_lo = (!s->planar && fr[2] == me) ? FACE_PERIODIC : FACE_NONE;
    B.pad_y_hi = (!s->planar && fr[3] == me) ? FACE_PERIODIC : FACE_NONE;
    B.bc_x_lo = m[0]; B.bc_x_hi = m[1]; B.bc_y_lo = m[2]; B.bc_y_hi = m[3]; B.bc_z_lo = m[4]; B.bc_z_hi = m[5];
    B.xwall = s->d_wall[0]; B.ywall = s->d_wall[1]; B.zwall = s->d_wall[2];
    B.planar = s->planar;
}

/* Peer-memory halos: every rank allocates a receive arena (two parity slots of x / y face buffers + flags), the CUDA IPC
 * handles travel through one ncclAllGather, and each rank maps the arenas of its face neighbours.  All ranks agree (MIN
 * all-reduce) on whether the mapping worked; otherwise, or with SARASMG_HALO=nccl, the grouped ncclSend/ncclRecv path stays. */
static int setup_ipc_halo(S *s) {
    const sarasmg_desc &d = s->d;
    const int np = d.n_proc;
    const LevelView &v0 = s->lev[0].v;
    const char *mode = getenv("SARASMG_HALO");
    const bool want = !(mode && strcmp(mode, "nccl") == 0);
    s->ipc_xc = (size_t)(s->planar ? 1 : v0.Fy) * v0.Fz;
    s->ipc_yc = s->planar ? 0 : (size_t)v0.Fx * v0.Fz;
    s->ipc_slot = 2 * s->ipc_xc + 2 * s->ipc_yc;
    const size_t arena_bytes = (2 * s->ipc_slot + 64) * sizeof(double);
    int ok = want ? 1 : 0;
    /* the halo push of the Jacobi sweeps additionally needs the neighbours' level slabs (SARASMG_NO_PUSH: keep the exchange kernels) */
    int okp = (want && !s->planar && !getenv("SARASMG_NO_PUSH")) ? 1 : 0;
    cudaIpcMemHandle_t mine, mine_slab;
    memset(&mine, 0, sizeof(mine));
    memset(&mine_slab, 0, sizeof(mine_slab));
    if (ok) {
        if (cudaMalloc(&s->ipc_arena, arena_bytes) != cudaSuccess) { cudaGetLastError(); s->ipc_arena = nullptr; ok = 0; }
        else {
            CU(cudaMemsetAsync(s->ipc_arena, 0, arena_bytes, s->st));
            if (cudaIpcGetMemHandle(&mine, s->ipc_arena) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
    }
    if (ok && okp && cudaIpcGetMemHandle(&mine_slab, s->slab) != cudaSuccess) { cudaGetLastError(); okp = 0; }
    if (!ok) okp = 0;
    /* all-gather {arena handle, slab handle, ok, ok_push} */
    const size_t rec = 2 * sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<unsigned char> hmine(rec, 0), hall(rec * np, 0);
    memcpy(hmine.data(), &mine, sizeof(mine));
    memcpy(hmine.data() + sizeof(mine), &mine_slab, sizeof(mine_slab));
    hmine[2 * sizeof(mine)] = (unsigned char)ok;
    hmine[2 * sizeof(mine) + 1] = (unsigned char)okp;
    unsigned char *dm = nullptr, *da = nullptr;
    CU(cudaMalloc(&dm, rec));
    CU(cudaMalloc(&da, rec * np));
    CU(cudaMemcpyAsync(dm, hmine.data(), rec, cudaMemcpyHostToDevice, s->st));
    NC(ncclAllGather(dm, da, rec, ncclUint8, s->comm, s->st));
    CU(cudaMemcpyAsync(hall.data(), da, rec * np, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (int r = 0; r < np; r++) { ok = ok && hall[r * rec + 2 * sizeof(mine)]; okp = okp && hall[r * rec + 2 * sizeof(mine) + 1]; }
    if (ok) {
        for (int f = 0; f < 4; f++) {
            const int r = d.face_ranks[f];
            if (r == SARASMG_PROC_NULL || r == d.rank || (s->planar && f >= 2)) continue;
            for (int g = 0; g < f; g++) if (d.face_ranks[g] == r && s->ipc_peer[g]) { s->ipc_peer[f] = s->ipc_peer[g]; s->slab_peer[f] = s->slab_peer[g]; }   /* same peer on two sides */
            if (s->ipc_peer[f]) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, hall.data() + r * rec, sizeof(h));
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            s->ipc_opened[s->ipc_nopened++] = ptr;
            s->ipc_peer[f] = (double *)ptr;
            if (okp) {
                memcpy(&h, hall.data() + r * rec + sizeof(h), sizeof(h));
                ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); okp = 0; }
                else { s->ipc_opened[s->ipc_nopened++] = ptr; s->slab_peer[f] = (double *)ptr; }
            }
        }
    }
    if (!ok) okp = 0;
    /* agree: MIN over ranks of the final status (reuses the small device buffers) */
    int *di = (int *)dm;
    int st2[2] = {ok, okp};
    CU(cudaMemcpyAsync(di, st2, 2 * sizeof(int), cudaMemcpyHostToDevice, s->st));
    NC(ncclAllReduce(di, di, 2, ncclInt, ncclMin, s->comm, s->st));
    CU(cudaMemcpyAsync(st2, di, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    ok = st2[0]; okp = st2[1];
    CU(cudaFree(dm)); CU(cudaFree(da));
    s->ipc_ok = (ok != 0);
    s->push_ok = (ok != 0 && okp != 0);
    if (!s->ipc_ok && want && d.rank == 0 && getenv("SARASMG_VERBOSE"))
        fprintf(stderr, "saras-mg: CUDA IPC halo buffers unavailable, using ncclSend/ncclRecv\n");
    return 0;
}

/* One-time exchange for the agglomerated coarsest solve: every rank learns every rank's face neighbours, wall flags and
 * coarsest-level x / y metric arrays (the z arrays are common: z is never split, parallel.cc:89-92). */
static int setup_agglomeration(S *s) {
    const sarasmg_desc &d = s->d;
    const int np = d.n_proc;
    if (np > 16 || s->D == 0) return 0;                 /* table size of the kernel; D == 0 would need every rank's wall planes */
    Level &L = s->lev[s->D];
    const LevelView &v = L.v;
    if (np == 1) {                                      /* single rank: a one-entry table over the level's own arrays */
        int m1[6];
        bc_modes(s, m1);
        CoarseBlock cb;
        for (int f = 0; f < 4; f++) { cb.nbr[f] = (d.face_ranks[f] == d.rank) ? 0 : -1; cb.wall[f] = (m1[f] != FACE_NONE) ? 1 : 0; }
        if (s->planar) { cb.nbr[2] = cb.nbr[3] = -1; cb.wall[2] = cb.wall[3] = 0; }
        cb.xixx = v.xixx; cb.xix2 = v.xix2; cb.etyy = v.etyy; cb.ety2 = v.ety2;
        CU(cudaMalloc(&s->agg_tab, sizeof(CoarseBlock)));
        CU(cudaMemcpyAsync(s->agg_tab, &cb, sizeof(CoarseBlock), cudaMemcpyHostToDevice, s->st));
        CU(cudaStreamSynchronize(s->st));
        s->agg_ok = true;
        return 0;
    }
    const int mx = v.Fx + 2, my = v.Fy + 2;
    const int rec = 8 + 2 * mx + 2 * my;                /* doubles per rank: 4 neighbours, 4 wall flags, xixx, xix2, etyy, ety2 */
    std::vector<double> mine(rec, 0.0), all((size_t)rec * np, 0.0);
    int m[6];
    bc_modes(s, m);                                     /* which x / y faces are physical walls does not depend on zeroBC */
    for (int f = 0; f < 4; f++) { mine[f] = d.face_ranks[f]; mine[4 + f] = (m[f] != FACE_NONE) ? 1.0 : 0.0; }
    CU(cudaMemcpyAsync(mine.data() + 8, L.met[0], mx * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(mine.data() + 8 + mx, L.met[1], mx * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(mine.data() + 8 + 2 * mx, L.met[2], my * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(mine.data() + 8 + 2 * mx + my, L.met[3], my * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    double *d_mine = nullptr, *d_all = nullptr;
    CU(cudaMalloc(&d_mine, rec * sizeof(double)));
    CU(cudaMalloc(&d_all, (size_t)rec * np * sizeof(double)));
    CU(cudaMemcpyAsync(d_mine, mine.data(), rec * sizeof(double), cudaMemcpyHostToDevice, s->st));
    NC(ncclAllGather(d_mine, d_all, (size_t)rec, ncclDouble, s->comm, s->st));
    CU(cudaMemcpyAsync(all.data(), d_all, (size_t)rec * np * sizeof(double), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaFree(d_mine)); CU(cudaFree(d_all));

    const int met_rec = 2 * mx + 2 * my;
    std::vector<double> met((size_t)met_rec * np);
    std::vector<CoarseBlock> tab(np);
    CU(cudaMalloc(&s->agg_met, met.size() * sizeof(double)));
    for (int r = 0; r < np; r++) {
        const double *a = all.data() + (size_t)r * rec;
        for (int f = 0; f < 4; f++) { tab[r].nbr[f] = (int)a[f]; tab[r].wall[f] = (int)a[4 + f]; }
        if (s->planar) { tab[r].nbr[2] = tab[r].nbr[3] = -1; tab[r].wall[2] = tab[r].wall[3] = 0; }
        memcpy(met.data() + (size_t)r * met_rec, a + 8, met_rec * sizeof(double));
        double *base = s->agg_met + (size_t)r * met_rec;
        tab[r].xixx = base + 1; tab[r].xix2 = base + mx + 1; tab[r].etyy = base + 2 * mx + 1; tab[r].ety2 = base + 2 * mx + my + 1;
    }
    CU(cudaMemcpyAsync(s->agg_met, met.data(), met.size() * sizeof(double), cudaMemcpyHostToDevice, s->st));
    CU(cudaMalloc(&s->agg_tab, np * sizeof(CoarseBlock)));
    CU(cudaMemcpyAsync(s->agg_tab, tab.data(), np * sizeof(CoarseBlock), cudaMemcpyHostToDevice, s->st));
    CU(cudaMalloc(&s->agg_P, (size_t)np * L.elems * sizeof(double)));
    CU(cudaMalloc(&s->agg_R, (size_t)np * L.elems * sizeof(double)));
    CU(cudaStreamSynchronize(s->st));
    s->agg_ok = true;
    return 0;
}

/* The single-CTA coarsest solve replays every rank block and runs the whole while(true) in one kernel: right for the few
 * hundred points a sensible vcDepth leaves, an unbounded single-kernel runtime for a shallow hierarchy (a 65^3 coarsest
 * level would keep one SM busy for up to maxCount ~ 3e5 sweeps).  Above this many points (all rank blocks together) the
 * host-driven loop with the multi-CTA sweep is used instead; it is not graph-capturable. */
static bool coarse_in_one_cta(const S *s) {
    const LevelView &v = s->lev[s->D].v;
    long long lim = 40000;
    if (const char *e = getenv("SARASMG_COARSE_CTA_MAX")) lim = atoll(e);
    return (long long)s->d.n_proc * v.nx * v.ny * v.nz <= lim;
}

/* solve (d3:190-247, d2:150-199) */
static int op_solve_coarsest(S *s, int *iters, bool sync) {
    const int l = s->vLevel;
    Level &L = s->lev[l];
    const sarasmg_desc &d = s->d;
    const int *fr = d.face_ranks;
    bool remote = false;
    for (int f = 0; f < 4; f++) if (fr[f] != SARASMG_PROC_NULL && fr[f] != d.rank) remote = true;
    ProfScope ps(s, SARASMG_K_COARSE_SOLVE, l);
    const bool one_cta = coarse_in_one_cta(s);
    if (s->agg_ok && !s->coarse_legacy && l == s->D && one_cta) {
        /* the whole while(true) of solve() in one CTA.  Several ranks: all-gather the coarsest arrays once and replay all
         * rank-blocks on every rank (k_coarse_solve_blocks / _smem); one rank: the same kernel over the level's own arrays */
        int m[6];
        bc_modes(s, m);
        const int wall = d.test_poisson ? (s->zeroBC ? FACE_NEG_MIRROR : FACE_WALL) : FACE_MIRROR;
        if (wall == FACE_WALL) return set_err("solve: non-homogeneous Dirichlet BC on level %d", l);
        const int nblk = d.n_proc;
        double *Pa = L.buf[SARASMG_PRESSURE];
        const double *Ra = L.buf[SARASMG_RESIDUAL];
        if (nblk > 1) {
            NC(ncclGroupStart());
            NC(ncclAllGather(L.buf[SARASMG_PRESSURE], s->agg_P, L.elems, ncclDouble, s->comm, s->st));
            NC(ncclAllGather(L.buf[SARASMG_RESIDUAL], s->agg_R, L.elems, ncclDouble, s->comm, s->st));
            NC(ncclGroupEnd());
            Pa = s->agg_P; Ra = s->agg_R;
        }
        const size_t smem = coarse_smem_bytes(L.v, d.dim, nblk);
        const int p2 = (divisors_pow2(s->planar, L.v.hx * L.v.hx, L.v.hx, L.v.hy, L.v.hz * L.v.hz, L.v.hz) && !s->no_pow2) ? 1 : 0;
        if (smem <= 200 * 1024 && !s->coarse_global) {
            if (s->planar) k_coarse_solve_smem<2><<<1, 256, smem, s->st>>>(Pa, Ra, (long long)L.elems, nblk, s->agg_tab, L.v, wall, m[4], m[5], p2, d.mg_tolerance, s->maxCount, s->d_iters);
            else k_coarse_solve_smem<3><<<1, 256, smem, s->st>>>(Pa, Ra, (long long)L.elems, nblk, s->agg_tab, L.v, wall, m[4], m[5], p2, d.mg_tolerance, s->maxCount, s->d_iters);
            TRY(kcheck(s, "k_coarse_solve_smem"));
        } else {
            if (s->planar) k_coarse_solve_blocks<2><<<1, 1024, 0, s->st>>>(Pa, Ra, (long long)L.elems, nblk, s->agg_tab, L.v, wall, m[4], m[5], d.mg_tolerance, s->maxCount, s->d_iters);
            else k_coarse_solve_blocks<3><<<1, 1024, 0, s->st>>>(Pa, Ra, (long long)L.elems, nblk, s->agg_tab, L.v, wall, m[4], m[5], d.mg_tolerance, s->maxCount, s->d_iters);
            TRY(kcheck(s, "k_coarse_solve_blocks"));
        }
        if (nblk > 1)
            CU(cudaMemcpyAsync(L.buf[SARASMG_PRESSURE], s->agg_P + (size_t)d.rank * L.elems, L.elems * sizeof(double), cudaMemcpyDeviceToDevice, s->st));
        CU(cudaMemcpyAsync(s->h_iters, s->d_iters, sizeof(int), cudaMemcpyDeviceToHost, s->st));
        if (sync) {
            CU(cudaStreamSynchronize(s->st));
            if (iters) *iters = *s->h_iters;
            s->last_coarse_iters = *s->h_iters;
        }
        return 0;
    }
    if ((d.n_proc == 1 || (!remote && !s->has_comm)) && one_cta) {
        BcPlan B;
        bc_plan(s, B);
        if ((B.bc_x_lo == FACE_WALL || B.bc_y_lo == FACE_WALL || B.bc_z_lo == FACE_WALL) && l != 0)
            return set_err("solve: non-homogeneous Dirichlet BC on level %d", l);
        if (s->planar) k_coarse_solve<2><<<1, 1024, 0, s->st>>>(L.buf[SARASMG_PRESSURE], L.buf[SARASMG_RESIDUAL], L.v, B, d.mg_tolerance, s->maxCount, s->d_iters);
        else k_coarse_solve<3><<<1, 1024, 0, s->st>>>(L.buf[SARASMG_PRESSURE], L.buf[SARASMG_RESIDUAL], L.v, B, d.mg_tolerance, s->maxCount, s->d_iters);
        TRY(kcheck(s, "k_coarse_solve"));
        CU(cudaMemcpyAsync(s->h_iters, s->d_iters, sizeof(int), cudaMemcpyDeviceToHost, s->st));
        if (sync) {
            CU(cudaStreamSynchronize(s->st));
            if (iters) *iters = *s->h_iters;
            s->last_coarse_iters = *s->h_iters;
        }
        return 0;
    }
    /* host-driven loop, one halo exchange and one MAX all-reduce per iteration (d3:195,233): the legacy multi-rank path and
     * the path of coarsest levels too large for one CTA */
    if (!one_cta && d.rank == 0 && getenv("SARASMG_VERBOSE"))
        fprintf(stderr, "saras-mg: coarsest level has %d x %d x %d points per rank: host-driven solve loop (consider a deeper V-cycle)\n", L.v.nx, L.v.ny, L.v.nz);
    if (s->capturing) return set_err("solve: the host-driven coarsest solve cannot be captured into a CUDA graph");
    int iterCount = 0;
    const long long cnt = (long long)L.v.nx * L.v.ny * L.v.nz;
    const unsigned nb = std::min(cdiv(cnt, 256), 1024u);
    TRY(ensure_part(s, nb));
    while (true) {
        TRY(impose_bc(s, l));
        TRY(gs_sweep(s, l));
        if (s->planar) k_resid_max<2><<<nb, 256, 0, s->st>>>(L.buf[SARASMG_PRESSURE], L.buf[SARASMG_RESIDUAL], L.v, s->d_part);
        else k_resid_max<3><<<nb, 256, 0, s->st>>>(L.buf[SARASMG_PRESSURE], L.buf[SARASMG_RESIDUAL], L.v, s->d_part);
        TRY(kcheck(s, "k_resid_max"));
        k_reduce_final<<<1, 1024, 0, s->st>>>(s->d_part, (int)nb, 1, 1u, s->d_scal);
        TRY(kcheck(s, "k_reduce_final"));
        std::vector<double> g1;
        TRY(gather_scalars(s, 1, g1));
        double globalMax = g1[0];
        for (int r = 1; r < d.n_proc; r++) globalMax = std::max(globalMax, g1[r]);
        if (globalMax < d.mg_tolerance) break;
        iterCount += 1;
        if (iterCount > s->maxCount) {
            if (d.print_residual && d.rank == 0) printf("WARNING: Iterations for solution at coarsest level not converging.\n");
            break;
        }
    }
    TRY(impose_bc(s, l));
    if (iters) *iters = iterCount;
    s->last_coarse_iters = iterCount;
    *s->h_iters = iterCount;
    return 0;
}

/* vCycle (p:219-280) */
static int op_vcycle_body(S *s);
static int op_vcycle(S *s, int *coarse_iters) {
    (void)coarse_iters;
    s->push_on = s->push_ok;            /* halo push of the Jacobi sweeps: inside V-cycles only (see push_level_ok) */
    const int rc = op_vcycle_body(s);
    s->push_on = false;
    memset(s->level_zero, 0, sizeof(s->level_zero));
    return rc;
}
/* after V-cycles that pushed halos: one full imposeBC on level 0 restores the edge pads of the array mgSolve hands back */
static int finalize_pads(S *s) {
    if (!s->pads_stale) return 0;
    s->pads_stale = false;
    const int lv = s->vLevel, zb = s->zeroBC;
    s->vLevel = 0; s->zeroBC = 0;
    const int rc = impose_bc(s, 0);
    s->vLevel = lv; s->zeroBC = zb;
    return rc;
}
static int op_vcycle_body(S *s) {
    const int D = s->D;
    s->vLevel = 0;
    s->zeroBC = 0;
    TRY(op_smooth(s, s->d.pre_smooth));
    s->zeroBC = 1;
    for (int i = 0; i < D; i++) {
        /* p:252 smoothedPres = pressureData: kept in place (see header comment) */
        if (fuse_level_ok(s, s->vLevel)) TRY(op_residual_coarsen_fused(s));
        else {
            TRY(op_compute_residual(s));
            TRY(op_coarsen(s));
        }
        {   /* p:258 pressureData(vLevel) = 0 */
            Level &L = s->lev[s->vLevel];
            ProfScope ps(s, SARASMG_K_MISC, s->vLevel);
            CU(cudaMemsetAsync(L.buf[SARASMG_PRESSURE], 0, L.elems * sizeof(double), s->st));
            if (s->vLevel < 32 && s->vLevel != D) s->level_zero[s->vLevel] = true;
        }
        if (s->vLevel == D) TRY(op_solve_coarsest(s, nullptr, false));
        else TRY(op_smooth(s, s->d.pre_smooth));
    }
    for (int i = 0; i < D; i++) {
        TRY(op_prolong(s));
        s->zeroBC = (s->vLevel == 0) ? 0 : 1;
        TRY(op_smooth(s, s->d.post_smooth));
    }
    return 0;
}

/* after a synchronisation: did a peer-memory wait time out? */
static int check_ipc_error(S *s) {
    if (!s->ipc_ok) return 0;
    unsigned e = 0;
    CU(cudaMemcpyAsync(&e, s->d_ipc_base + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    if (e) return set_err("peer-memory halo exchange %u timed out on rank %d: a neighbour never signalled (rank lost, or ranks out of step)", e & 0x7fffffffu, s->d.rank);
    return 0;
}

/* One V-cycle of mgSolve.  The cycle is a fixed sequence of several hundred launches, most of them on latency-bound
 * coarse levels, so it is captured once into a CUDA graph and replayed: kernels, memsets, the peer-memory exchanges (their
 * exchange numbers come from a device word, see IpcPlan), the NCCL all-gathers of the agglomerated coarsest solve and the
 * iteration-count read-back.  A graph is keyed by the level-buffer assignment (Jacobi ping-pong) and the profiling setup. */
static unsigned graph_prof_key(const S *s) {
    const unsigned sm = s->d.gs_smooth ? 0x40000000u : 0u;      /* sarasmg_set_smoother: one graph per smoother */
    if (!(s->d.flags & (SARASMG_FLAG_PROFILE | SARASMG_FLAG_TIME_RUN))) return sm;
    return sm + 1u + (unsigned)std::min(s->prof_max_level, 1000) + ((s->d.flags & SARASMG_FLAG_TIME_RUN) ? 100000u : 0u);
}
static int run_vcycle(S *s) {
    /* several ranks: the graph holds the NCCL all-gathers of the agglomerated coarsest solve and must be destroyed before
     * the communicator (sarasmg_destroy) */
    const bool capturable = s->use_graph && !s->gs_legacy && s->D > 0 &&    /* D == 0: the cycle is one smooth(); nothing to replay */
                            coarse_in_one_cta(s) && (s->d.n_proc == 1 || (s->graph_multi && s->agg_ok && !s->coarse_legacy && s->ipc_ok));   /* the ncclSend/ncclRecv halo fallback stays eager */
    if (!capturable) return op_vcycle(s, nullptr);
    const double *key = s->lev[0].buf[SARASMG_PRESSURE];
    const unsigned pkey = graph_prof_key(s);
    GraphSlot *g = nullptr;
    for (GraphSlot *q : s->graphs) if (q->key == key && q->prof_key == pkey) g = q;
    if (!g) {
        g = new GraphSlot();
        g->key = key; g->prof_key = pkey; g->exec = nullptr;
        const long long l0 = s->launches;
        const bool stale_before = s->pads_stale;
        s->pads_stale = false;
        s->capturing = true; s->cap_sites = 0; s->cap_prof = &g->prof;
        cudaError_t ce = cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal);
        int rc = (ce == cudaSuccess) ? op_vcycle(s, nullptr) : 1;
        cudaGraph_t graph = nullptr;
        cudaError_t ce2 = cudaStreamEndCapture(s->st, &graph);
        s->capturing = false;
        if (ce != cudaSuccess || ce2 != cudaSuccess || rc) {
            if (graph) cudaGraphDestroy(graph);
            delete g;
            if (rc && ce == cudaSuccess) return rc;
            return set_err("CUDA graph capture of the V-cycle failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ce2));
        }
        ce = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { delete g; return set_err("cudaGraphInstantiate: %s", cudaGetErrorString(ce)); }
        g->launches = s->launches - l0;
        s->launches = l0;
        g->sites = s->cap_sites;
        g->stale_pads = s->pads_stale;          /* set by the captured op_smooth on level 0: a replay has to set it again */
        s->pads_stale = stale_before;
        s->graphs.push_back(g);
    } else {
        /* the capture performed the host-side state changes of one cycle; a replay has to repeat them */
        if (!s->d.gs_smooth && ((s->d.pre_smooth + s->d.post_smooth) & 1))
            for (int l = 0; l < s->D; l++) {
                Level &L = s->lev[l];
                std::swap(L.buf[SARASMG_PRESSURE], L.buf[SARASMG_TMP]);
                std::swap(L.off[SARASMG_PRESSURE], L.off[SARASMG_TMP]);
                std::swap(L.tmapP[SARASMG_PRESSURE], L.tmapP[SARASMG_TMP]);
                std::swap(L.tmapR[SARASMG_PRESSURE], L.tmapR[SARASMG_TMP]);
                std::swap(L.tmapP16[SARASMG_PRESSURE], L.tmapP16[SARASMG_TMP]);
                std::swap(L.tmapR16[SARASMG_PRESSURE], L.tmapR16[SARASMG_TMP]);
            }
        s->vLevel = 0;
        s->zeroBC = 0;
    }
    if (s->last_graph) { CU(cudaStreamSynchronize(s->st)); TRY(graph_prof_fold(s)); }
    if (g->sites) {
        k_set_u32<<<1, 1, 0, s->st>>>(s->d_ipc_base, s->ipc_epoch);
        TRY(kcheck(s, "k_set_u32"));
        s->ipc_base_h = s->ipc_epoch;
    }
    CU(cudaGraphLaunch(g->exec, s->st));
    if (g->stale_pads) s->pads_stale = true;
    s->ipc_epoch += g->sites;
    s->launches += g->launches;
    if (!g->prof.empty()) s->last_graph = g;
    return 0;
}

/* everything mgSolve does between the host copies (p:105-148), from the device input buffers */
static int solve_core(S *s, const double *in_p, const double *in_rhs, int *cycles_done, double *hist, int *citers) {
    const sarasmg_desc &d = s->d;
    Level &L0 = s->lev[0];
    s->vLevel = 0;
    /* p:107-115: zero the hierarchy, load the cores of rhs / inFn into level 0 (one fused pass over level 0) */
    for (int l = 1; l <= s->D; l++) {
        Level &L = s->lev[l];
        CU(cudaMemsetAsync(L.buf[SARASMG_PRESSURE], 0, L.elems * sizeof(double), s->st));
        CU(cudaMemsetAsync(L.buf[SARASMG_RESIDUAL], 0, L.elems * sizeof(double), s->st));
    }
    const unsigned nbp = std::min(cdiv((long long)L0.v.Fx * L0.v.Fy, 2), 148u * 16u);
    const long long shell = 2LL * L0.v.Fy * L0.v.Fz + 2LL * L0.v.nx * L0.v.Fz + 2LL * L0.v.nx * L0.v.ny;
    const unsigned nbs = std::min(cdiv(shell, 256), 1024u);
    TRY(ensure_part(s, 2 * (size_t)(nbp + nbs) + 4096));
    {
        ProfScope ps(s, SARASMG_K_MISC, 0);
        /* compensated partial sums of the rhs core: part[2b], part[2b+1], b < nbp */
        k_load_inputs<<<nbp, 256, 0, s->st>>>(in_p, in_rhs, L0.buf[SARASMG_PRESSURE], L0.buf[SARASMG_RESIDUAL], L0.v, s->d_part);
        TRY(kcheck(s, "k_load_inputs"));
    }
    TRY(update_pads(s, 0, L0.buf[SARASMG_RESIDUAL]));       /* p:117 */
    TRY(update_pads(s, 0, L0.buf[SARASMG_PRESSURE]));       /* p:118 */

    if (!d.test_poisson) {                                  /* p:120-129 */
        ProfScope ps(s, SARASMG_K_MISC, 0);
        /* the pads' partial sums follow the core's, then ONE compensated reduction of all of them: d_scal[0] is the
         * correctly rounded sum of the whole padded array (blitz::mean, p:122, is that sum / count) */
        k_reduce_shell<<<nbs, 256, 0, s->st>>>(L0.buf[SARASMG_RESIDUAL], L0.v, s->d_part + 2 * (size_t)nbp);
        TRY(kcheck(s, "k_reduce_shell"));
        k_reduce_final_cs<<<1, 1024, 0, s->st>>>(s->d_part, (int)(nbp + nbs), s->d_scal);
        TRY(kcheck(s, "k_reduce_final_cs"));
        std::vector<double> sums;
        TRY(gather_scalars(s, 1, sums));
        const double cnt = (double)((long long)L0.v.Fx * L0.v.Fy * L0.v.Fz);
        double globalAvg = 0.0;
        for (int r = 0; r < d.n_proc; r++) globalAvg += sums[r] / cnt;      /* sum of local means, rank order (MPI_SUM, p:125) */
        globalAvg /= d.n_proc;
        if (s->mean_override) globalAvg = s->mean_value;
        s->last_mean = globalAvg;
        k_sub_rows<<<nbp, 256, 0, s->st>>>(L0.buf[SARASMG_RESIDUAL], L0.v, globalAvg);
        TRY(kcheck(s, "k_sub_rows"));
    }

    double prevResidual = 0.0;
    int cycles = 0;
    CU(cudaEventRecord(s->ev_solve0, s->st));
    for (int i = 0; i < d.vc_count; i++) {                  /* p:132-148 */
        *s->h_iters = 0;
        TRY(run_vcycle(s));
        double mgResidual = 0.0;
        TRY(op_compute_error(s, d.res_type, &mgResidual));
        if (d.res_type == 0) CU(cudaStreamSynchronize(s->st));
        TRY(graph_prof_fold(s));                            /* the stream is idle here: computeError read its scalars back */
        s->last_coarse_iters = *s->h_iters;
        if (s->last_coarse_iters > s->maxCount && coarse_in_one_cta(s) && (d.n_proc == 1 || (s->agg_ok && !s->coarse_legacy)) && d.print_residual && d.rank == 0)
            printf("WARNING: Iterations for solution at coarsest level not converging.\n");    /* d3:240-243 */
        if (hist) hist[i] = mgResidual;
        if (citers) citers[i] = s->last_coarse_iters;
        cycles = i + 1;
        if (d.print_residual && d.rank == 0) {
            if (d.test_poisson) printf("\nResidual after V Cycle %d is %.16g\n", i, mgResidual);
            else printf("\nResidual after V Cycle %d is %g\n", i, mgResidual);
        }
        if (!d.test_poisson) {
            if (std::fabs(prevResidual - mgResidual) < d.mg_tolerance) break;
            prevResidual = mgResidual;
        }
    }
    TRY(finalize_pads(s));
    CU(cudaEventRecord(s->ev_solve1, s->st));
    CU(cudaEventSynchronize(s->ev_solve1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s->ev_solve0, s->ev_solve1));
    s->last_solve_ms = ms;
    if (cycles_done) *cycles_done = cycles;
    TRY(check_ipc_error(s));
    return 0;
}

static int analytic_deviation(S *s, double *out) {
    Level &L0 = s->lev[0];
    const sarasmg_desc &d = s->d;
    const int ubx = L0.v.nx - 1, uby = L0.v.ny - 1;
    const long long cnt = (long long)L0.v.nx * L0.v.ny * L0.v.nz;
    const unsigned nb = std::min(cdiv(cnt, 256), 2048u);
    TRY(ensure_part(s, nb));
    k_analytic_dev<<<nb, 256, 0, s->st>>>(L0.buf[SARASMG_PRESSURE], L0.v, d.dim, d.x_rank * ubx, d.y_rank * uby,
                                          ubx * d.npx / 2, uby * d.npy / 2, s->d_part);
    TRY(kcheck(s, "k_analytic_dev"));
    k_reduce_final<<<1, 1024, 0, s->st>>>(s->d_part, (int)nb, 1, 1u, s->d_scal);
    TRY(kcheck(s, "k_reduce_final"));
    std::vector<double> g;
    TRY(gather_scalars(s, 1, g));
    double m = g[0];
    for (int r = 1; r < d.n_proc; r++) m = std::max(m, g[r]);
    *out = m;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* extern "C"                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
extern "C" {

const char *sarasmg_last_error(void) { return g_err.c_str(); }
int sarasmg_version(void) { return 100; }

int sarasmg_device_count(int *count) {
    if (!count) return set_err("null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return set_err("cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = n;
    return 0;
}

int sarasmg_get_unique_id(void *out) {
    static_assert(sizeof(ncclUniqueId) == SARASMG_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return 0;
}

/* builds everything into *s; on failure the caller tears the partially built solver down */
static int create_impl(const sarasmg_desc *desc, S *s) {
    s->d = *desc;
    s->planar = (desc->dim == 2);
    s->D = desc->vc_depth;
    s->vLevel = 0; s->zeroBC = 0;
    s->d_part = nullptr; s->part_cap = 0;
    s->d_in_p = s->d_in_rhs = nullptr;
    s->has_comm = false;
    s->launches = 0;
    s->t_comp = s->t_tran = 0.0; s->bucket = 0;
    s->last_solve_ms = 0.0; s->last_coarse_iters = 0;
    s->mean_override = 0; s->mean_value = 0.0; s->last_mean = 0.0;
    s->prof_max_level = 1 << 20;
    s->gs_legacy = getenv("SARASMG_GS_LEGACY") != nullptr;
    s->no_pow2 = getenv("SARASMG_NO_POW2") != nullptr;
    s->prolong_legacy = getenv("SARASMG_PROLONG_LEGACY") != nullptr;
    s->coarsen_legacy = getenv("SARASMG_COARSEN_LEGACY") != nullptr;
    s->no_fuse = getenv("SARASMG_NO_FUSE") != nullptr;
    s->no_zero_skip = getenv("SARASMG_NO_ZEROSKIP") != nullptr;
    s->fuse_min_points = getenv("SARASMG_FUSE_MIN") ? atof(getenv("SARASMG_FUSE_MIN")) : (getenv("SARASMG_TMA_MIN") ? 0.0 : 4.0e8);   /* tests that force the TMA kernels onto small grids get the fused pass too */
    for (int q = 0; q < 8; q++) s->user_ev[q] = nullptr;
    s->proj_ok = false; s->d_presP = nullptr;
    for (int q = 0; q < 3; q++) { s->d_vel[q] = nullptr; s->vel_elems[q] = 0; }
    for (int q = 0; q < 6; q++) s->d_pmet[q] = nullptr;
    s->d_psync[0] = s->d_psync[1] = nullptr;
    s->impl_ok = false;
    for (int q = 0; q < 6; q++) s->d_impl_met[q] = nullptr;
    for (int q = 0; q < 3; q++) { s->d_nse[q] = nullptr; s->d_guess[q] = nullptr; }
    memset(s->prof_ms, 0, sizeof(s->prof_ms));
    memset(s->prof_n, 0, sizeof(s->prof_n));
    for (int q = 0; q < 3; q++) s->d_wall[q] = nullptr;
    for (int q = 0; q < 2; q++) s->xs[q] = s->xr[q] = s->ys[q] = s->yr[q] = nullptr;
    s->d_gather = nullptr;
    s->use_graph = getenv("SARASMG_NO_GRAPH") == nullptr;
    s->graph_multi = !(getenv("SARASMG_GRAPH_MULTI") && atoi(getenv("SARASMG_GRAPH_MULTI")) == 0);   /* default on; SARASMG_GRAPH_MULTI=0: eager launches with several ranks */
    s->capturing = false; s->cap_sites = 0; s->cap_prof = nullptr; s->last_graph = nullptr;
    s->d_ipc_base = nullptr; s->ipc_base_h = 0;
    s->ipc_cta_max = getenv("SARASMG_IPC_CTA_MAX") ? atoll(getenv("SARASMG_IPC_CTA_MAX")) : 4096;
    s->ipc_ok = false; s->ipc_arena = nullptr; s->ipc_nopened = 0; s->ipc_epoch = 0;
    for (int q = 0; q < 4; q++) { s->ipc_peer[q] = nullptr; s->slab_peer[q] = nullptr; }
    for (int q = 0; q < 8; q++) s->ipc_opened[q] = nullptr;
    s->slab = nullptr; s->slab_elems = 0; s->push_ok = false; s->push_on = false; s->pads_stale = false;
    memset(s->push_add, 0, sizeof(s->push_add));
    memset(s->level_zero, 0, sizeof(s->level_zero));
    s->agg_P = s->agg_R = nullptr; s->agg_tab = nullptr; s->agg_met = nullptr; s->agg_ok = false;
    s->coarse_legacy = getenv("SARASMG_COARSE_LEGACY") != nullptr;
    s->coarse_global = getenv("SARASMG_COARSE_GLOBAL") != nullptr;
    CU(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    CU(cudaEventCreate(&s->ev_solve0));
    CU(cudaEventCreate(&s->ev_solve1));

    const int D = s->D;
    s->lev.resize(D + 1);
    size_t moff[3] = {0, 0, 0};
    {   /* one slab for the three arrays of every level (sizes as computed per level below) */
        size_t tot = 0;
        for (int l = 0; l <= D; l++) {
            long long F[3];
            for (int dd = 0; dd < 3; dd++) F[dd] = (s->planar && dd == 1) ? 3 : (((desc->n0[dd] - 1) >> l) + 1) + 2;
            const long long pitch = ((F[2] - 2 + 6 + 3) / 4) * 4;
            const size_t elems = (((size_t)F[0] * F[1] * pitch + 16 + 31) / 32) * 32;
            tot += 3 * elems;
        }
        s->slab_elems = tot;
        CU(cudaMalloc(&s->slab, tot * sizeof(double)));
        CU(cudaMemsetAsync(s->slab, 0, tot * sizeof(double), s->st));
    }
    size_t slab_used = 0;
    for (int l = 0; l <= D; l++) {
        Level &L = s->lev[l];
        LevelView &v = L.v;
        int n[3];
        for (int dd = 0; dd < 3; dd++) {
            const int n0 = desc->n0[dd];
            if (s->planar && dd == 1) { n[dd] = 1; continue; }
            if (((n0 - 1) >> l) << l != (n0 - 1) || ((n0 - 1) >> l) < 2)
                return set_err("sarasmg_create: level-0 size %d along axis %d cannot be coarsened %d times (need >= 3 points, parser.cc:307-337)", n0, dd, l);
            n[dd] = ((n0 - 1) >> l) + 1;
        }
        v.nx = n[0]; v.ny = n[1]; v.nz = n[2];
        v.Fx = n[0] + 2; v.Fy = n[1] + 2; v.Fz = n[2] + 2;
        v.pitch = ((v.nz + 6 + 3) / 4) * 4;
        v.plane = (long long)v.Fy * v.pitch;
        L.elems = (((size_t)v.Fx * v.plane + 16 + 31) / 32) * 32;      /* multiple of 256 bytes: every array starts 256-byte aligned */
        for (int a = 0; a < 3; a++) {
            if (slab_used + L.elems > s->slab_elems) return set_err("sarasmg_create: level slab layout mismatch");
            L.off[a] = slab_used;
            L.buf[a] = s->slab + slab_used;
            slab_used += L.elems;
        }
        v.hx = desc->hx ? desc->hx[l] : 0.0;   v.hy = desc->hy ? desc->hy[l] : 0.0;   v.hz = desc->hz ? desc->hz[l] : 0.0;
        v.hx2 = desc->hx2 ? desc->hx2[l] : 0.0; v.hz2 = desc->hz2 ? desc->hz2[l] : 0.0; v.hzhx = desc->hzhx ? desc->hzhx[l] : 0.0;
        v.hxhy = desc->hxhy ? desc->hxhy[l] : 0.0; v.hyhz = desc->hyhz ? desc->hyhz[l] : 0.0; v.hxhyhz = desc->hxhyhz ? desc->hxhyhz[l] : 0.0;
        const double *src[6] = {desc->xixx, desc->xix2, desc->etyy, desc->ety2, desc->ztzz, desc->ztz2};
        const int Fd[3] = {v.Fx, v.Fy, v.Fz};
        for (int m = 0; m < 6; m++) {
            const int dd = m / 2;
            CU(cudaMalloc(&L.met[m], (size_t)(Fd[dd] + 2) * sizeof(double)));
            CU(cudaMemsetAsync(L.met[m], 0, (size_t)(Fd[dd] + 2) * sizeof(double), s->st));
            if (src[m] && !(s->planar && dd == 1))
                CU(cudaMemcpyAsync(L.met[m], src[m] + moff[dd], (size_t)Fd[dd] * sizeof(double), cudaMemcpyHostToDevice, s->st));
        }
        L.x_uniform = false;
        if (desc->xix2) {
            const double *a = desc->xix2 + moff[0] + 1;     /* core entries 0..nx-1 */
            L.x_uniform = true;
            for (int i = 1; i < v.nx; i++) if (a[i] != a[0]) { L.x_uniform = false; break; }
        }
        L.z_uniform = false;
        if (desc->ztz2) {
            const double *a = desc->ztz2 + moff[2] + 1;
            L.z_uniform = true;
            for (int k = 1; k < v.nz; k++) if (a[k] != a[0]) { L.z_uniform = false; break; }
        }
        L.zz_uniform = L.z_uniform && desc->ztzz;           /* ... and so has ztzz: the strip Gauss-Seidel kernel then needs no z tables */
        if (L.zz_uniform) {
            const double *a = desc->ztzz + moff[2] + 1;
            for (int k = 1; k < v.nz; k++) if (a[k] != a[0]) { L.zz_uniform = false; break; }
        }
        L.gs_tasks = nullptr; L.gs_sync = nullptr; L.gs_smem = 0;
        L.gss_tasks = nullptr; L.gss_sync = nullptr; L.gss_rden = nullptr; L.gss_xpair = nullptr; L.gss_smem = 0; L.gss_ok = false;
        for (int dd = 0; dd < 3; dd++) moff[dd] += (size_t)Fd[dd];
        v.xixx = L.met[0] + 1; v.xix2 = L.met[1] + 1; v.etyy = L.met[2] + 1; v.ety2 = L.met[3] + 1; v.ztzz = L.met[4] + 1; v.ztz2 = L.met[5] + 1;
    }
    for (int l = 0; l <= D; l++) TRY(build_gs_plan(s, s->lev[l]));
    for (int l = 0; l <= D; l++) TRY(build_gs_strip_plan(s, s->lev[l]));
    for (int l = 0; l <= D; l++) TRY(make_level_tmaps(s, s->lev[l]));
    CU(cudaFuncSetAttribute(k_coarse_solve_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_coarse_solve_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_stencil_tma<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_stencil_tma<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_stencil_tma_push<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_stencil_tma_push<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_resid_coarsen_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_F_SMEM));
    CU(cudaFuncSetAttribute(k_resid_coarsen_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_F_SMEM));
    CU(cudaFuncSetAttribute(k_stencil_tma<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_stencil_tma<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_stencil_tma<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_stencil_tma<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_TMA_SMEM));
    CU(cudaFuncSetAttribute(k_gs_skew<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_gs_skew<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    {
        int dev = 0, smem_max = 227 * 1024;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        CU(cudaFuncSetAttribute(k_gs_strip<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 1024));
        CU(cudaFuncSetAttribute(k_gs_strip<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 1024));
    }
    {   /* p:360-361 maxCount */
        const LevelView &v = s->lev[D].v;
        s->maxCount = (v.Fx - 1) * (v.Fy - 1) * (v.Fz - 1);
    }
    {   /* Dirichlet planes */
        const LevelView &v0 = s->lev[0].v;
        size_t wn[3];
        if (!s->planar) { wn[0] = (size_t)v0.Fy * v0.Fz; wn[1] = (size_t)v0.Fx * v0.Fz; wn[2] = (size_t)v0.Fx * v0.Fy; }
        else { wn[0] = (size_t)v0.Fz; wn[1] = 0; wn[2] = (size_t)v0.Fx; }
        const double *ws[3] = {desc->x_wall, desc->y_wall, desc->z_wall};
        for (int q = 0; q < 3; q++) {
            if (desc->test_poisson && wn[q]) {
                if (!ws[q]) return set_err("sarasmg_create: test_poisson needs wall array %d", q);
                CU(cudaMalloc(&s->d_wall[q], wn[q] * sizeof(double)));
                CU(cudaMemcpyAsync(s->d_wall[q], ws[q], wn[q] * sizeof(double), cudaMemcpyHostToDevice, s->st));
            }
        }
    }
    CU(cudaMalloc(&s->d_ipc_base, 64));
    CU(cudaMemsetAsync(s->d_ipc_base, 0, 64, s->st));
    CU(cudaMalloc(&s->d_scal, 64 * sizeof(double)));
    CU(cudaMallocHost(&s->h_scal, (size_t)std::max(64, 8 * desc->n_proc) * sizeof(double)));
    CU(cudaMalloc(&s->d_iters, sizeof(int)));
    CU(cudaMemsetAsync(s->d_iters, 0, sizeof(int), s->st));
    CU(cudaMallocHost(&s->h_iters, sizeof(int)));
    *s->h_iters = 0;
    TRY(ensure_part(s, 1 << 16));

    if (desc->n_proc > 1) {
        if (!desc->nccl_id) return set_err("sarasmg_create: n_proc > 1 needs nccl_id");
        ncclUniqueId id;
        memcpy(&id, desc->nccl_id, sizeof(id));
        NC(ncclCommInitRank(&s->comm, desc->n_proc, id, desc->rank));
        s->has_comm = true;
        const LevelView &v0 = s->lev[0].v;
        const size_t xc = (size_t)v0.Fy * v0.Fz, yc = (size_t)v0.Fx * v0.Fz;
        for (int q = 0; q < 2; q++) {
            CU(cudaMalloc(&s->xs[q], xc * sizeof(double))); CU(cudaMalloc(&s->xr[q], xc * sizeof(double)));
            CU(cudaMalloc(&s->ys[q], yc * sizeof(double))); CU(cudaMalloc(&s->yr[q], yc * sizeof(double)));
        }
        CU(cudaMalloc(&s->d_gather, (size_t)desc->n_proc * 8 * sizeof(double)));
        TRY(setup_agglomeration(s));
        TRY(setup_ipc_halo(s));
    }
    if (desc->n_proc == 1) TRY(setup_agglomeration(s));
    s->d.nccl_id = nullptr;     /* the caller's buffers are not retained */
    s->d.hx = s->d.hy = s->d.hz = s->d.hx2 = s->d.hz2 = s->d.hzhx = s->d.hxhy = s->d.hyhz = s->d.hxhyhz = nullptr;
    s->d.xixx = s->d.xix2 = s->d.etyy = s->d.ety2 = s->d.ztzz = s->d.ztz2 = nullptr;
    s->d.x_wall = s->d.y_wall = s->d.z_wall = nullptr;
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

int sarasmg_create(const sarasmg_desc *desc, sarasmg_handle *out) {
    if (!desc || !out) return set_err("sarasmg_create: null argument");
    if (desc->dim != 2 && desc->dim != 3) return set_err("sarasmg_create: dim must be 2 or 3");
    if (desc->vc_depth < 0 || desc->vc_depth > 30) return set_err("sarasmg_create: bad vc_depth");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return set_err("sarasmg_create: no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(ce));
    CU(cudaSetDevice(desc->device));
    S *s = new S();
    const int rc = create_impl(desc, s);
    if (rc) {                               /* free whatever was allocated; keep the message of the failure */
        const std::string why = g_err;
        s->d.flags &= ~SARASMG_FLAG_TIME_RUN;
        sarasmg_destroy(s);
        cudaGetLastError();
        g_err = why;
        return rc;
    }
    *out = s;
    return 0;
}

int sarasmg_destroy(sarasmg_handle s) {
    if (!s) return 0;
    cudaSetDevice(s->d.device);
    cudaStreamSynchronize(s->st);
    /* graphs first: a captured NCCL collective keeps a reference on its communicator until the graph is gone */
    for (GraphSlot *g : s->graphs) {
        for (auto &e : g->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        cudaGraphExecDestroy(g->exec);
        delete g;
    }
    s->graphs.clear();
    if (s->has_comm) ncclCommDestroy(s->comm);
    for (auto &L : s->lev) {
        for (int m = 0; m < 6; m++) cudaFree(L.met[m]);
        cudaFree(L.gs_tasks); cudaFree(L.gs_sync);
        cudaFree(L.gss_tasks); cudaFree(L.gss_sync); cudaFree(L.gss_rden); cudaFree(L.gss_xpair);
    }
    for (int q = 0; q < 3; q++) if (s->d_wall[q]) cudaFree(s->d_wall[q]);
    for (int q = 0; q < 2; q++) { cudaFree(s->xs[q]); cudaFree(s->xr[q]); cudaFree(s->ys[q]); cudaFree(s->yr[q]); }
    cudaFree(s->d_gather);
    cudaFree(s->d_ipc_base);
    for (int q = 0; q < s->ipc_nopened; q++) cudaIpcCloseMemHandle(s->ipc_opened[q]);
    cudaFree(s->ipc_arena);
    cudaFree(s->slab);
    cudaFree(s->agg_P); cudaFree(s->agg_R); cudaFree(s->agg_tab); cudaFree(s->agg_met);
    cudaFree(s->d_part); cudaFree(s->d_scal); cudaFree(s->d_iters);
    cudaFreeHost(s->h_scal); cudaFreeHost(s->h_iters);
    cudaFree(s->d_in_p); cudaFree(s->d_in_rhs);
    for (int q = 0; q < 3; q++) cudaFree(s->d_vel[q]);
    for (int q = 0; q < 6; q++) cudaFree(s->d_pmet[q]);
    cudaFree(s->d_psync[0]); cudaFree(s->d_psync[1]);
    for (int q = 0; q < 6; q++) cudaFree(s->d_impl_met[q]);
    for (int q = 0; q < 3; q++) { cudaFree(s->d_nse[q]); cudaFree(s->d_guess[q]); }
    cudaFree(s->d_presP);
    for (auto &e : s->pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto &e : s->ev_pool) cudaEventDestroy(e);
    cudaEventDestroy(s->ev_solve0); cudaEventDestroy(s->ev_solve1);
    for (int q = 0; q < 8; q++) if (s->user_ev[q]) cudaEventDestroy(s->user_ev[q]);
    for (auto &pr : s->pinned) cudaHostUnregister(const_cast<void *>(pr.first));
    cudaStreamDestroy(s->st);
    if ((s->d.flags & SARASMG_FLAG_TIME_RUN) && s->d.rank == 0) {
        /* poisson::~poisson under -DTIME_RUN (poisson.cc:645-652) */
        printf("%-50s%.6f\n", "Time taken in computation within smooth: ", s->t_comp);
        printf("%-50s%.6f\n", "Time taken in data-transfer within smooth: ", s->t_tran);
    }
    delete s;
    return 0;
}

static int ensure_inputs(S *s) {
    Level &L0 = s->lev[0];
    if (!s->d_in_p) {
        CU(cudaMalloc(&s->d_in_p, L0.elems * sizeof(double)));
        CU(cudaMalloc(&s->d_in_rhs, L0.elems * sizeof(double)));
        CU(cudaMemsetAsync(s->d_in_p, 0, L0.elems * sizeof(double), s->st));
        CU(cudaMemsetAsync(s->d_in_rhs, 0, L0.elems * sizeof(double), s->st));
    }
    return 0;
}

int sarasmg_upload(sarasmg_handle s, const double *p_in, const double *rhs) {
    if (!s || !p_in || !rhs) return set_err("sarasmg_upload: null argument");
    CU(cudaSetDevice(s->d.device));
    TRY(ensure_inputs(s));
    TRY(copy_h2d(s, s->lev[0], s->d_in_p, p_in));
    TRY(copy_h2d(s, s->lev[0], s->d_in_rhs, rhs));
    return 0;
}

int sarasmg_solve_resident(sarasmg_handle s, int *cycles_done, double *residual_history, int *coarse_iters) {
    if (!s) return set_err("sarasmg_solve_resident: null handle");
    if (!s->d_in_p) return set_err("sarasmg_solve_resident: call sarasmg_upload first");
    CU(cudaSetDevice(s->d.device));
    return solve_core(s, s->d_in_p, s->d_in_rhs, cycles_done, residual_history, coarse_iters);
}

int sarasmg_download(sarasmg_handle s, double *p_out) {
    if (!s || !p_out) return set_err("sarasmg_download: null argument");
    CU(cudaSetDevice(s->d.device));
    TRY(copy_d2h(s, s->lev[0], p_out, s->lev[0].buf[SARASMG_PRESSURE]));      /* p:151: whole padded box */
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

/* Page-locking of caller memory happens ONLY on request (sarasmg_host_register): the library cannot know how long a caller's
 * buffer lives, and a registration that outlives its buffer is undefined behaviour.  The caller must unregister (or destroy
 * the handle) before freeing the range. */
int sarasmg_host_register(sarasmg_handle s, const void *ptr, unsigned long long bytes) {
    if (!s || !ptr) return set_err("sarasmg_host_register: null argument");
    CU(cudaSetDevice(s->d.device));
    for (auto &pr : s->pinned) if (pr.first == ptr && pr.second >= (size_t)bytes) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return 0;   /* already pinned by the caller */
    cudaGetLastError();
    cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), (size_t)bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err("sarasmg_host_register: %s", cudaGetErrorString(e)); }
    s->pinned.push_back({ptr, (size_t)bytes});
    return 0;
}

int sarasmg_host_unregister(sarasmg_handle s, const void *ptr) {
    if (!s || !ptr) return set_err("sarasmg_host_unregister: null argument");
    CU(cudaSetDevice(s->d.device));
    for (size_t q = 0; q < s->pinned.size(); q++)
        if (s->pinned[q].first == ptr) {
            CU(cudaStreamSynchronize(s->st));
            cudaHostUnregister(const_cast<void *>(ptr));
            cudaGetLastError();
            s->pinned.erase(s->pinned.begin() + q);
            return 0;
        }
    return 0;       /* not registered through this handle: nothing to do */
}

int sarasmg_solve(sarasmg_handle s, double *p_inout, const double *rhs, int *cycles_done, double *residual_history, int *coarse_iters) {
    if (!s || !p_inout || !rhs) return set_err("sarasmg_solve: null argument");
    CU(cudaSetDevice(s->d.device));
    TRY(sarasmg_upload(s, p_inout, rhs));
    TRY(sarasmg_solve_resident(s, cycles_done, residual_history, coarse_iters));
    TRY(sarasmg_download(s, p_inout));
    return 0;
}

/* ---- device-resident projection step ---- */
int sarasmg_proj_setup(sarasmg_handle s, const sarasmg_proj_desc *d) {
    if (!s || !d) return set_err("sarasmg_proj_setup: null argument");
    if (s->planar) return set_err("sarasmg_proj_setup: the device-resident projection step is built for multigrid_d3 only");
    if (!d->xi_x_staggr || !d->et_y_staggr || !d->zt_z_staggr || !d->xi_x_colloc || !d->et_y_colloc || !d->zt_z_colloc)
        return set_err("sarasmg_proj_setup: null metric array");
    CU(cudaSetDevice(s->d.device));
    TRY(ensure_inputs(s));
    const LevelView &v0 = s->lev[0].v;
    ProjView &V = s->pv;
    V.nx = v0.nx; V.ny = v0.ny; V.nz = v0.nz;
    V.dXi = d->d_xi; V.dEt = d->d_et; V.dZt = d->d_zt;
    V.p2 = (is_pow2(V.dXi) && is_pow2(V.dEt) && is_pow2(V.dZt) && !s->no_pow2) ? 1 : 0;
    V.rXi = 1.0 / V.dXi; V.rEt = 1.0 / V.dEt; V.rZt = 1.0 / V.dZt;
    const double *src[6] = {d->xi_x_staggr, d->et_y_staggr, d->zt_z_staggr, d->xi_x_colloc, d->et_y_colloc, d->zt_z_colloc};
    const int len[6] = {v0.nx + 2, v0.ny + 2, v0.nz + 2, v0.nx + 1, v0.ny + 1, v0.nz + 1};
    for (int q = 0; q < 6; q++) {
        if (!s->d_pmet[q]) CU(cudaMalloc(&s->d_pmet[q], sizeof(double) * len[q]));
        CU(cudaMemcpyAsync(s->d_pmet[q], src[q], sizeof(double) * len[q], cudaMemcpyHostToDevice, s->st));
    }
    V.xs = s->d_pmet[0] + 1; V.ys = s->d_pmet[1] + 1; V.zs = s->d_pmet[2] + 1;
    V.xc = s->d_pmet[3] + 1; V.yc = s->d_pmet[4] + 1; V.zc = s->d_pmet[5] + 1;
    const size_t ve[3] = {(size_t)(v0.nx + 1) * (v0.ny + 2) * (v0.nz + 2), (size_t)(v0.nx + 2) * (v0.ny + 1) * (v0.nz + 2),
                          (size_t)(v0.nx + 2) * (v0.ny + 2) * (v0.nz + 1)};
    for (int q = 0; q < 3; q++) {
        s->vel_elems[q] = ve[q];
        if (!s->d_vel[q]) CU(cudaMalloc(&s->d_vel[q], sizeof(double) * ve[q]));
        CU(cudaMemsetAsync(s->d_vel[q], 0, sizeof(double) * ve[q], s->st));
    }
    {   /* Pp.syncData() between ranks: face / edge partners and the message buffers */
        const sarasmg_desc &dd = s->d;
        const int me = dd.rank;
        auto find_rank = [&](int xr, int yr) { return ((yr + dd.npy) % dd.npy) * dd.npx + ((xr + dd.npx) % dd.npx); };   /* parallel.cc:133-160 */
        s->edge_ranks[0] = find_rank(dd.x_rank - 1, dd.y_rank - 1); s->edge_ranks[1] = find_rank(dd.x_rank - 1, dd.y_rank + 1);
        s->edge_ranks[2] = find_rank(dd.x_rank + 1, dd.y_rank - 1); s->edge_ranks[3] = find_rank(dd.x_rank + 1, dd.y_rank + 1);
        const long long nxf = (long long)v0.ny * v0.nz, nyf = (long long)v0.nx * v0.nz;
        for (int f = 0; f < 4; f++) {
            s->psync.face_remote[f] = (dd.face_ranks[f] != SARASMG_PROC_NULL && dd.face_ranks[f] != me) ? 1 : 0;
            s->psync.off_face[f] = (f < 2) ? f * nxf : 2 * nxf + (f - 2) * nyf;
            s->psync.off_edge[f] = 2 * nxf + 2 * nyf + (long long)f * v0.nz;
        }
        if (dd.n_proc > 1) {
            const size_t tot = (size_t)(2 * nxf + 2 * nyf + 4LL * v0.nz);
            for (int q = 0; q < 2; q++) if (!s->d_psync[q]) CU(cudaMalloc(&s->d_psync[q], tot * sizeof(double)));
        }
    }
    if (!s->d_presP) CU(cudaMalloc(&s->d_presP, s->lev[0].elems * sizeof(double)));
    CU(cudaMemsetAsync(s->d_presP, 0, s->lev[0].elems * sizeof(double), s->st));
    CU(cudaStreamSynchronize(s->st));
    s->proj_ok = true;
    return 0;
}

/* which: 0 Vx, 1 Vy, 2 Vz (compact Blitz layouts), 3 P, 4 Pp (the initial guess of the next solve): padded staggered arrays */
int sarasmg_proj_upload(sarasmg_handle s, int which, const double *host) {
    if (!s || !host || which < 0 || which > 4) return set_err("sarasmg_proj_upload: bad argument");
    if (!s->proj_ok) return set_err("sarasmg_proj_upload: call sarasmg_proj_setup first");
    CU(cudaSetDevice(s->d.device));
    if (which < 3) CU(cudaMemcpyAsync(s->d_vel[which], host, sizeof(double) * s->vel_elems[which], cudaMemcpyHostToDevice, s->st));
    else TRY(copy_h2d(s, s->lev[0], which == 3 ? s->d_presP : s->d_in_p, host));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}
/* which: 0..4 as above, 5 mgRHS (the divergence / dt the last sarasmg_proj_divergence wrote) */
int sarasmg_proj_download(sarasmg_handle s, int which, double *host) {
    if (!s || !host || which < 0 || which > 5) return set_err("sarasmg_proj_download: bad argument");
    if (!s->proj_ok) return set_err("sarasmg_proj_download: call sarasmg_proj_setup first");
    CU(cudaSetDevice(s->d.device));
    if (which < 3) CU(cudaMemcpyAsync(host, s->d_vel[which], sizeof(double) * s->vel_elems[which], cudaMemcpyDeviceToHost, s->st));
    else TRY(copy_d2h(s, s->lev[0], host, which == 3 ? s->d_presP : (which == 4 ? s->d_in_p : s->d_in_rhs)));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

static int proj_divergence(S *s, double inv_dt) {
    const LevelView &v0 = s->lev[0].v;
    const dim3 grid(cdiv((long long)v0.Fy * v0.Fz, 128), v0.Fx, 1);
    k_proj_divergence<<<grid, 128, 0, s->st>>>(s->d_vel[0], s->d_vel[1], s->d_vel[2], s->d_in_rhs, v0, s->pv, inv_dt);
    return kcheck(s, "k_proj_divergence");
}
static int proj_correct(S *s, double dt) {
    Level &L0 = s->lev[0];
    const LevelView &v0 = L0.v;
    /* inFn.F = pressureData(0) over the whole padded box (p:151): Pp, which is also the initial guess of the next solve */
    k_copy_padded<<<std::min(cdiv((long long)v0.Fx * v0.Fy, 1), 148u * 8u), 256, 0, s->st>>>(L0.buf[SARASMG_PRESSURE], s->d_in_p, v0);
    TRY(kcheck(s, "k_copy_padded"));
    /* Pp.syncData() (hydro_d3.cc:347): faces with faceRanks, edges with edgeRanks */
    const int xper = (s->d.face_ranks[0] == s->d.rank) ? 1 : 0, yper = (s->d.face_ranks[2] == s->d.rank) ? 1 : 0;
    if (xper || yper) {         /* this rank is its own neighbour along a periodic, unsplit axis */
        k_proj_sync_faces<<<std::min(cdiv(2LL * v0.ny * v0.nz + 2LL * v0.nx * v0.nz, 256), 2048u), 256, 0, s->st>>>(s->d_in_p, v0, xper, yper);
        TRY(kcheck(s, "k_proj_sync_faces"));
    }
    if (s->d.n_proc == 1) {
        k_proj_sync_edges<<<cdiv(4LL * v0.nz, 256), 256, 0, s->st>>>(s->d_in_p, v0);
        TRY(kcheck(s, "k_proj_sync_edges"));
    } else {
        if (!s->has_comm) return set_err("sarasmg_proj_correct: several ranks but no communicator");
        const long long nxf = (long long)v0.ny * v0.nz, nyf = (long long)v0.nx * v0.nz;
        const unsigned nb = std::min(cdiv(2 * nxf + 2 * nyf + 4LL * v0.nz, 256), 2048u);
        k_proj_sync_pack<<<nb, 256, 0, s->st>>>(s->d_in_p, v0, s->psync, s->d_psync[0]);
        TRY(kcheck(s, "k_proj_sync_pack"));
        const int *fr = s->d.face_ranks, *er = s->edge_ranks;
        const ProjSyncPlan &Q = s->psync;
        double *sb = s->d_psync[0], *rb = s->d_psync[1];
        const size_t fc[4] = {(size_t)nxf, (size_t)nxf, (size_t)nyf, (size_t)nyf};
        NC(ncclGroupStart());
        /* pairs (send one side, receive the opposite pad) keep messages matched when one peer is the neighbour on both sides */
        for (int a = 0; a < 4; a += 2) {
            if (Q.face_remote[a])     NC(ncclSend(sb + Q.off_face[a],     fc[a], ncclDouble, fr[a],     s->comm, s->st));
            if (Q.face_remote[a + 1]) NC(ncclRecv(rb + Q.off_face[a + 1], fc[a], ncclDouble, fr[a + 1], s->comm, s->st));
            if (Q.face_remote[a + 1]) NC(ncclSend(sb + Q.off_face[a + 1], fc[a], ncclDouble, fr[a + 1], s->comm, s->st));
            if (Q.face_remote[a])     NC(ncclRecv(rb + Q.off_face[a],     fc[a], ncclDouble, fr[a],     s->comm, s->st));
        }
        /* edges: my (x lo, y lo) corner column lands in the (x hi, y hi) pad of edgeRanks(0), ... (mpidata.cc:296-305) */
        const int pair[4][2] = {{0, 3}, {3, 0}, {1, 2}, {2, 1}};
        for (int q = 0; q < 4; q++) {
            NC(ncclSend(sb + Q.off_edge[pair[q][0]], (size_t)v0.nz, ncclDouble, er[pair[q][0]], s->comm, s->st));
            NC(ncclRecv(rb + Q.off_edge[pair[q][1]], (size_t)v0.nz, ncclDouble, er[pair[q][1]], s->comm, s->st));
        }
        NC(ncclGroupEnd());
        k_proj_sync_unpack<<<nb, 256, 0, s->st>>>(s->d_in_p, v0, s->psync, s->d_psync[1]);
        TRY(kcheck(s, "k_proj_sync_unpack"));
    }
    const dim3 grid(cdiv((long long)v0.Fy * v0.Fz, 128), v0.Fx, 1);
    k_proj_correct<<<grid, 128, 0, s->st>>>(s->d_in_p, s->d_presP, s->d_vel[0], s->d_vel[1], s->d_vel[2], v0, s->pv, dt);
    return kcheck(s, "k_proj_correct");
}

int sarasmg_proj_divergence(sarasmg_handle s, double inv_dt) {
    if (!s || !s->proj_ok) return set_err("sarasmg_proj_divergence: call sarasmg_proj_setup first");
    CU(cudaSetDevice(s->d.device));
    TRY(proj_divergence(s, inv_dt));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}
int sarasmg_proj_correct(sarasmg_handle s, double dt) {
    if (!s || !s->proj_ok) return set_err("sarasmg_proj_correct: call sarasmg_proj_setup first");
    CU(cudaSetDevice(s->d.device));
    TRY(proj_correct(s, dt));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}
int sarasmg_proj_step(sarasmg_handle s, double dt, int *cycles_done, double *residual_history, int *coarse_iters) {
    if (!s || !s->proj_ok) return set_err("sarasmg_proj_step: call sarasmg_proj_setup first");
    CU(cudaSetDevice(s->d.device));
    TRY(proj_divergence(s, 1.0 / dt));
    TRY(solve_core(s, s->d_in_p, s->d_in_rhs, cycles_done, residual_history, coarse_iters));
    TRY(proj_correct(s, dt));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

/* ---- implicit velocity solves ---- */
static ImplGeom impl_geom(const S *s, int comp) {
    ImplGeom g;
    const LevelView &v0 = s->lev[0].v;
    const int n[3] = {v0.nx, v0.ny, v0.nz};
    const int per[3] = {s->d.x_per, s->d.y_per, s->d.z_per};
    for (int d = 0; d < 3; d++) {
        g.stag[d] = (d != comp) ? 1 : 0;
        g.cs[d] = g.stag[d] ? n[d] : n[d] - 1;
        g.ext[d] = g.cs[d] + 2;
        const int cut = (g.stag[d] && !per[d]) ? 1 : 0;                /* field.cc:206-219 on one rank */
        g.lb[d] = cut; g.ub[d] = g.cs[d] - 1 - cut;
    }
    g.mx = s->d_impl_met[g.stag[0] ? 3 : 0] + 1; g.my = s->d_impl_met[g.stag[1] ? 4 : 1] + 1; g.mz = s->d_impl_met[g.stag[2] ? 5 : 2] + 1;
    return g;
}
static ImplCoef impl_coef(const S *s) {
    const sarasmg_impl_desc &d = s->impl;
    ImplCoef c;
    c.hy2hz2 = d.hy2hz2; c.hz2hx2 = d.hz2hx2; c.hx2hy2 = d.hx2hy2; c.hx2hy2hz2 = d.hx2hy2hz2;
    c.hx = d.hx; c.hy = d.hy; c.hz = d.hz; c.dt = d.dt; c.Re = d.Re; c.inverseRe = 1.0 / d.Re;     /* hydro.cc:65 */
    c.den1 = d.hx2hy2hz2 * 2.0 * d.Re; c.rden1 = 1.0 / c.den1;
    c.den2 = d.Re * d.hx2hy2hz2; c.rden2 = 1.0 / c.den2;
    c.hxx = d.hx * d.hx; c.hyy = d.hy * d.hy; c.hzz = d.hz * d.hz;
    c.rhxx = 1.0 / c.hxx; c.rhyy = 1.0 / c.hyy; c.rhzz = 1.0 / c.hzz;
    c.p2 = (is_pow2(c.hxx) && is_pow2(c.hyy) && is_pow2(c.hzz) && !s->no_pow2) ? 1 : 0;
    return c;
}
int sarasmg_impl_setup(sarasmg_handle s, const sarasmg_impl_desc *d) {
    if (!s || !d) return set_err("sarasmg_impl_setup: null argument");
    if (!s->proj_ok) return set_err("sarasmg_impl_setup: call sarasmg_proj_setup first (the velocity components live there)");
    if (s->d.n_proc != 1) return set_err("sarasmg_impl_setup: one rank only for now");
    if (!d->xix2_colloc || !d->ety2_colloc || !d->ztz2_colloc || !d->xix2_staggr || !d->ety2_staggr || !d->ztz2_staggr)
        return set_err("sarasmg_impl_setup: null metric array");
    if (!(d->Re > 0.0) || !(d->dt > 0.0)) return set_err("sarasmg_impl_setup: dt and Re must be positive");
    CU(cudaSetDevice(s->d.device));
    const LevelView &v0 = s->lev[0].v;
    const double *src[6] = {d->xix2_colloc, d->ety2_colloc, d->ztz2_colloc, d->xix2_staggr, d->ety2_staggr, d->ztz2_staggr};
    const int len[6] = {v0.nx + 1, v0.ny + 1, v0.nz + 1, v0.nx + 2, v0.ny + 2, v0.nz + 2};
    for (int q = 0; q < 6; q++) {
        if (!s->d_impl_met[q]) CU(cudaMalloc(&s->d_impl_met[q], sizeof(double) * len[q]));
        CU(cudaMemcpyAsync(s->d_impl_met[q], src[q], sizeof(double) * len[q], cudaMemcpyHostToDevice, s->st));
    }
    /* nseRHS / guessedVelocity of a component are allocated when the component is first used (impl_component) */
    CU(cudaStreamSynchronize(s->st));
    s->impl = *d;
    s->impl.xix2_colloc = s->impl.ety2_colloc = s->impl.ztz2_colloc = s->impl.xix2_staggr = s->impl.ety2_staggr = s->impl.ztz2_staggr = nullptr;
    s->impl_ok = true;
    return 0;
}
static int impl_component(S *s, int comp) {
    if (!s->d_nse[comp]) {
        CU(cudaMalloc(&s->d_nse[comp], sizeof(double) * s->vel_elems[comp]));
        CU(cudaMemsetAsync(s->d_nse[comp], 0, sizeof(double) * s->vel_elems[comp], s->st));
    }
    if (!s->d_guess[comp]) CU(cudaMalloc(&s->d_guess[comp], sizeof(double) * s->vel_elems[comp]));
    return 0;
}
int sarasmg_impl_upload_rhs(sarasmg_handle s, int comp, const double *host) {
    if (!s || !host || comp < 0 || comp > 2) return set_err("sarasmg_impl_upload_rhs: bad argument");
    if (!s->impl_ok) return set_err("sarasmg_impl_upload_rhs: call sarasmg_impl_setup first");
    CU(cudaSetDevice(s->d.device));
    TRY(impl_component(s, comp));
    CU(cudaMemcpyAsync(s->d_nse[comp], host, sizeof(double) * s->vel_elems[comp], cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}
/* imposeUBCs / VBCs / WBCs (hydro.cc:386-445) on one rank */
static int impl_bcs(S *s, int comp, double *F) {
    const ImplGeom g = impl_geom(s, comp);
    const int xper = s->d.x_per ? 1 : 0, yper = s->d.y_per ? 1 : 0;
    if (xper || yper) {
        const long long cnt = (xper ? (long long)g.cs[1] * g.cs[2] : 0) + (yper ? (long long)g.cs[0] * g.cs[2] : 0);
        k_impl_sync_faces<<<face_blocks(cnt), 256, 0, s->st>>>(F, g, xper, yper);
        TRY(kcheck(s, "k_impl_sync_faces"));
    }
    k_impl_sync_edges<<<cdiv(g.cs[2], 256), 256, 0, s->st>>>(F, g);
    TRY(kcheck(s, "k_impl_sync_edges"));
    const int order[6] = {0, 1, 2, 3, 5, 4};
    for (int q = 0; q < 6; q++) {
        const int w = order[q], ax = w / 2, hi = w & 1;
        const int t = s->impl.bc_type[comp][w];
        if (t == SARASMG_BC_NONE) continue;
        const int wall_lo = g.lb[ax] - 1, wall_hi = g.ub[ax] + 1;      /* field.cc:270-285 */
        const int dst = hi ? wall_hi : wall_lo;
        int src = dst + (hi ? -1 : 1);                                 /* shiftVal, boundary.cc:55-84 */
        if (t == SARASMG_BC_PERIODIC_CC) src = hi ? wall_lo + 2 : wall_hi - 2;
        if (t == SARASMG_BC_PERIODIC_FC) src = hi ? wall_lo + 1 : wall_hi - 1;
        const int a1 = (ax == 0) ? 1 : 0, a2 = (ax == 2) ? 1 : 2;
        k_impl_plane<<<face_blocks((long long)g.ext[a1] * g.ext[a2]), 256, 0, s->st>>>(F, g, ax, dst, src, t, s->impl.bc_value[comp][w]);
        TRY(kcheck(s, "k_impl_plane"));
    }
    return 0;
}
int sarasmg_impl_bcs(sarasmg_handle s, int comp) {
    if (!s || comp < 0 || comp > 2) return set_err("sarasmg_impl_bcs: bad argument");
    if (!s->impl_ok) return set_err("sarasmg_impl_bcs: call sarasmg_impl_setup first");
    CU(cudaSetDevice(s->d.device));
    TRY(impl_bcs(s, comp, s->d_vel[comp]));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}
int sarasmg_impl_solve(sarasmg_handle s, int comp, int *sweeps, double *max_error) {
    if (!s || comp < 0 || comp > 2) return set_err("sarasmg_impl_solve: bad argument");
    if (!s->impl_ok) return set_err("sarasmg_impl_solve: call sarasmg_impl_setup first");
    CU(cudaSetDevice(s->d.device));
    TRY(impl_component(s, comp));
    const ImplGeom g = impl_geom(s, comp);
    const ImplCoef c = impl_coef(s);
    const dim3 gs(cdiv((long long)g.ext[1] * g.ext[2], 256), g.ext[0], 1);
    const int bn[3] = {g.ub[0] - g.lb[0] + 1, g.ub[1] - g.lb[1] + 1, g.ub[2] - g.lb[2] + 1};
    if (bn[0] < 1 || bn[1] < 1 || bn[2] < 1) return set_err("sarasmg_impl_solve: empty bulk");
    const dim3 ge(cdiv((long long)bn[1] * bn[2], 256), bn[0], 1);
    const size_t nb = (size_t)ge.x * ge.y * ge.z;
    TRY(ensure_part(s, nb + 16));
    long long iterCount = 0;
    double maxError = 0.0;
    for (;;) {
        k_impl_sweep<<<gs, 256, 0, s->st>>>(s->d_vel[comp], s->d_nse[comp], s->d_guess[comp], g, c);
        TRY(kcheck(s, "k_impl_sweep"));
        std::swap(s->d_vel[comp], s->d_guess[comp]);                    /* V.Vx.F = guessedVelocity.Vx (every entry of it was just written) */
        TRY(impl_bcs(s, comp, s->d_vel[comp]));
        k_impl_error<<<ge, 256, 0, s->st>>>(s->d_vel[comp], s->d_nse[comp], g, c, s->d_part);
        TRY(kcheck(s, "k_impl_error"));
        k_reduce_final<<<1, 1024, 0, s->st>>>(s->d_part, (int)nb, 1, 1u, s->d_scal);
        TRY(kcheck(s, "k_reduce_final"));
        CU(cudaMemcpyAsync(s->h_scal, s->d_scal, sizeof(double), cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        maxError = std::max(s->h_scal[0], 0.0);                         /* vxMax runs over the whole array, whose non-bulk entries are 0 */
        if (maxError < s->impl.cn_tolerance) break;
        iterCount += 1;
        if (iterCount > s->impl.max_iterations) {
            if (max_error) *max_error = maxError;
            return set_err("sarasmg_impl_solve: Jacobi iterations for solution of V%c not converging (the reference aborts here, hydro_d3.cc:415-421)", "xyz"[comp]);
        }
    }
    if (sweeps) *sweeps = (int)(iterCount + 1);
    if (max_error) *max_error = maxError;
    return 0;
}

int sarasmg_analytic_deviation(sarasmg_handle s, double *max_dev) {
    if (!s || !max_dev) return set_err("sarasmg_analytic_deviation: null argument");
    CU(cudaSetDevice(s->d.device));
    return analytic_deviation(s, max_dev);
}

int sarasmg_level_dims(sarasmg_handle s, int level, int n[3], int F[3]) {
    if (!s || level < 0 || level > s->D) return set_err("sarasmg_level_dims: bad level");
    const LevelView &v = s->lev[level].v;
    n[0] = v.nx; n[1] = v.ny; n[2] = v.nz; F[0] = v.Fx; F[1] = v.Fy; F[2] = v.Fz;
    return 0;
}

int sarasmg_level_set(sarasmg_handle s, int level, int which, const double *host) {
    if (!s || level < 0 || level > s->D || which < 0 || which > 2 || !host) return set_err("sarasmg_level_set: bad argument");
    CU(cudaSetDevice(s->d.device));
    TRY(copy_h2d(s, s->lev[level], s->lev[level].buf[which], host));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

int sarasmg_level_get(sarasmg_handle s, int level, int which, double *host) {
    if (!s || level < 0 || level > s->D || which < 0 || which > 2 || !host) return set_err("sarasmg_level_get: bad argument");
    CU(cudaSetDevice(s->d.device));
    TRY(copy_d2h(s, s->lev[level], host, s->lev[level].buf[which]));
    CU(cudaStreamSynchronize(s->st));
    return 0;
}

int sarasmg_set_level(sarasmg_handle s, int v) {
    if (!s || v < 0 || v > s->D) return set_err("sarasmg_set_level: bad level");
    s->vLevel = v;
    return 0;
}
int sarasmg_set_zero_bc(sarasmg_handle s, int z) {
    if (!s) return set_err("null handle");
    s->zeroBC = z ? 1 : 0;
    return 0;
}

#define ENTER(s)                                         \
    if (!(s)) return set_err("null handle");             \
    CU(cudaSetDevice((s)->d.device));

int sarasmg_op_update_pads(sarasmg_handle s, int which) {
    ENTER(s);
    if (which < 0 || which > 2) return set_err("bad array id");
    TRY(update_pads(s, s->vLevel, s->lev[s->vLevel].buf[which]));
    CU(cudaStreamSynchronize(s->st));
    return check_ipc_error(s);
}
int sarasmg_op_impose_bc(sarasmg_handle s) { ENTER(s); TRY(impose_bc(s, s->vLevel)); CU(cudaStreamSynchronize(s->st)); return check_ipc_error(s); }
int sarasmg_op_smooth(sarasmg_handle s, int count) { ENTER(s); TRY(op_smooth(s, count)); CU(cudaStreamSynchronize(s->st)); return check_ipc_error(s); }
int sarasmg_op_compute_residual(sarasmg_handle s) { ENTER(s); TRY(op_compute_residual(s)); CU(cudaStreamSynchronize(s->st)); return check_ipc_error(s); }
int sarasmg_op_coarsen(sarasmg_handle s) { ENTER(s); TRY(op_coarsen(s)); CU(cudaStreamSynchronize(s->st)); return 0; }
int sarasmg_op_prolong(sarasmg_handle s) { ENTER(s); TRY(op_prolong(s)); CU(cudaStreamSynchronize(s->st)); return 0; }
int sarasmg_op_solve_coarsest(sarasmg_handle s, int *iters) { ENTER(s); return op_solve_coarsest(s, iters, true); }
int sarasmg_op_compute_error(sarasmg_handle s, int norm, double *value) {
    ENTER(s);
    if (!value) return set_err("null value");
    return op_compute_error(s, norm, value);
}
int sarasmg_op_vcycle(sarasmg_handle s, int *coarse_iters) {
    ENTER(s);
    *s->h_iters = 0;
    TRY(op_vcycle(s, nullptr));
    TRY(finalize_pads(s));
    CU(cudaStreamSynchronize(s->st));
    s->last_coarse_iters = *s->h_iters;
    if (coarse_iters) *coarse_iters = s->last_coarse_iters;
    return 0;
}

/* ---- self tests ---- */

/* testProlong (d3:689-719, d2:481-507): level 1 holds the linear pattern (rank+1)*1000 + 100 i + 10 j + k sampled at the
 * even fine points, it is prolonged and compared with the same pattern on level 0 (trilinear interpolation is exact for
 * linear data).  The reference's version is stale: it fills the coarse level with COARSE indices and then indexes
 * residualData(vLevel - 1) = residualData(-1) (d3:716); this is the test as intended.  The fine array starts at 0, so
 * the fused "+= smoothedPres" adds nothing. */
int sarasmg_selftest_prolong(sarasmg_handle s, double *max_err) {
    ENTER(s);
    if (s->D < 1) return set_err("selftest_prolong needs vc_depth >= 1");
    std::vector<double> expect, coarse, got;
    for (int l = 0; l <= 1; l++) {
        const LevelView &v = s->lev[l].v;
        std::vector<double> &A = (l == 0) ? expect : coarse;
        A.assign((size_t)v.Fx * v.Fy * v.Fz, 0.0);
        for (int i = 0; i < v.nx; i++) for (int j = 0; j < v.ny; j++) for (int k = 0; k < v.nz; k++) {
            const size_t o = ((size_t)(i + 1) * v.Fy + (j + 1)) * v.Fz + (k + 1);
            const int st = 1 << l;      /* level-1 point (i,j,k) is fine point (2i,2j,2k) */
            A[o] = s->planar ? (s->d.rank + 1) * 100 + i * st * 10 + k * st
                             : (s->d.rank + 1) * 1000 + i * st * 100 + j * st * 10 + k * st;
        }
    }
    got.assign(expect.size(), 0.0);
    TRY(sarasmg_level_set(s, 0, SARASMG_PRESSURE, got.data()));
    TRY(sarasmg_level_set(s, 1, SARASMG_PRESSURE, coarse.data()));
    s->vLevel = 1;
    TRY(op_prolong(s));
    TRY(sarasmg_level_get(s, 0, SARASMG_PRESSURE, got.data()));
    double m = 0.0;
    for (size_t q = 0; q < got.size(); q++) m = std::max(m, std::fabs(got[q] - expect[q]));
    *max_err = m;
    return 0;
}

/* testTransfer (d3:722-777, d2:510-550): fill the core of every level with (rank+1)*1000 + 100 i + 10 j + k,
 * run updatePads on every level and compare the x/y pads with what the neighbour must have sent.
 * (The reference fills only level 0 and indexes one level past the hierarchy, d3:759-764; this is the test as intended.) */
int sarasmg_selftest_transfer(sarasmg_handle s, double *max_err) {
    ENTER(s);
    double worst = 0.0;
    const int *fr = s->d.face_ranks;
    for (int l = 0; l <= s->D; l++) {
        const LevelView &v = s->lev[l].v;
        std::vector<double> A((size_t)v.Fx * v.Fy * v.Fz, 0.0);
        auto at = [&](int i, int j, int k) -> double & { return A[((size_t)(i + 1) * v.Fy + (j + 1)) * v.Fz + (k + 1)]; };
        auto val = [&](int rank, int i, int j, int k) -> double {
            return s->planar ? (rank + 1) * 100 + i * 10 + k : (rank + 1) * 1000 + i * 100 + j * 10 + k;
        };
        for (int i = 0; i < v.nx; i++) for (int j = 0; j < v.ny; j++) for (int k = 0; k < v.nz; k++) at(i, j, k) = val(s->d.rank, i, j, k);
        TRY(sarasmg_level_set(s, l, SARASMG_PRESSURE, A.data()));
        s->vLevel = l;
        TRY(update_pads(s, l, s->lev[l].buf[SARASMG_PRESSURE]));
        TRY(sarasmg_level_get(s, l, SARASMG_PRESSURE, A.data()));
        const int xE = v.nx - 1, yE = v.ny - 1;
        for (int j = 0; j < v.ny; j++) for (int k = 0; k < v.nz; k++) {
            if (fr[0] != SARASMG_PROC_NULL) worst = std::max(worst, std::fabs(at(-1, j, k) - val(fr[0], xE - 1, j, k)));
            if (fr[1] != SARASMG_PROC_NULL) worst = std::max(worst, std::fabs(at(xE + 1, j, k) - val(fr[1], 1, j, k)));
        }
        if (!s->planar)
            for (int i = 0; i < v.nx; i++) for (int k = 0; k < v.nz; k++) {
                if (fr[2] != SARASMG_PROC_NULL) worst = std::max(worst, std::fabs(at(i, -1, k) - val(fr[2], i, yE - 1, k)));
                if (fr[3] != SARASMG_PROC_NULL) worst = std::max(worst, std::fabs(at(i, yE + 1, k) - val(fr[3], i, 1, k)));
            }
    }
    s->vLevel = 0;
    *max_err = worst;
    return 0;
}

/* testPeriodic (d3:780-857, d2:553-602): sin*cos*cos sampled on every level, imposeBC, pads compared with the
 * analytic periodic continuation.  Coordinates are the rank's level-0 x/y/zStaggr arrays (length F_d(0), base -1). */
int sarasmg_selftest_periodic(sarasmg_handle s, const double *xs, const double *ys, const double *zs,
                              double x_len, double y_len, double z_len, double *max_err) {
    ENTER(s);
    if (!xs || !zs || (!s->planar && !ys)) return set_err("selftest_periodic: null coordinates");
    double worst = 0.0;
    const LevelView &v0 = s->lev[0].v;
    const int xE0 = v0.nx - 1, yE0 = v0.ny - 1, zE0 = v0.nz - 1;
    const double two_pi = 2.0 * M_PI;
    auto X = [&](int i) { return xs[i + 1]; };
    auto Y = [&](int j) { return s->planar ? 0.0 : ys[j + 1]; };
    auto Z = [&](int k) { return zs[k + 1]; };
    auto f = [&](double x, double y, double z) {
        return s->planar ? std::sin(two_pi * x / x_len) * std::cos(two_pi * z / z_len)
                         : std::sin(two_pi * x / x_len) * std::cos(two_pi * y / y_len) * std::cos(two_pi * z / z_len);
    };
    const int zb = s->zeroBC;
    for (int l = 0; l <= s->D; l++) {
        const LevelView &v = s->lev[l].v;
        const int st = 1 << l;
        std::vector<double> A((size_t)v.Fx * v.Fy * v.Fz, 0.0);
        auto at = [&](int i, int j, int k) -> double & { return A[((size_t)(i + 1) * v.Fy + (j + 1)) * v.Fz + (k + 1)]; };
        for (int i = 0; i < v.nx; i++) for (int j = 0; j < v.ny; j++) for (int k = 0; k < v.nz; k++) at(i, j, k) = f(X(i * st), Y(j * st), Z(k * st));
        TRY(sarasmg_level_set(s, l, SARASMG_PRESSURE, A.data()));
        s->vLevel = l;
        s->zeroBC = 1;
        TRY(impose_bc(s, l));
        TRY(sarasmg_level_get(s, l, SARASMG_PRESSURE, A.data()));
        const int xE = v.nx - 1, yE = v.ny - 1, zE = v.nz - 1;
        if (s->d.x_per) {
            const double xl = X(0) - (X(st) - X(0)), xr = X(xE0) + (X(xE0) - X(xE0 - st));
            for (int j = 0; j < v.ny; j++) for (int k = 0; k < v.nz; k++) {
                worst = std::max(worst, std::fabs(at(-1, j, k) - f(xl, Y(j * st), Z(k * st))));
                worst = std::max(worst, std::fabs(at(xE + 1, j, k) - f(xr, Y(j * st), Z(k * st))));
            }
        }
        if (!s->planar && s->d.y_per) {
            const double yl = Y(0) - (Y(st) - Y(0)), yr = Y(yE0) + (Y(yE0) - Y(yE0 - st));
            for (int i = 0; i < v.nx; i++) for (int k = 0; k < v.nz; k++) {
                worst = std::max(worst, std::fabs(at(i, -1, k) - f(X(i * st), yl, Z(k * st))));
                worst = std::max(worst, std::fabs(at(i, yE + 1, k) - f(X(i * st), yr, Z(k * st))));
            }
        }
        if (s->d.z_per) {
            const double zl = Z(0) - (Z(st) - Z(0)), zr = Z(zE0) + (Z(zE0) - Z(zE0 - st));
            for (int i = 0; i < v.nx; i++) for (int j = 0; j < v.ny; j++) {
                worst = std::max(worst, std::fabs(at(i, j, -1) - f(X(i * st), Y(j * st), zl)));
                worst = std::max(worst, std::fabs(at(i, j, zE + 1) - f(X(i * st), Y(j * st), zr)));
            }
        }
    }
    s->vLevel = 0;
    s->zeroBC = zb;
    *max_err = worst;
    return 0;
}

/* ---- instrumentation ---- */
int sarasmg_timers(sarasmg_handle s, double *comp, double *tran) {
    ENTER(s);
    TRY(prof_fold(s));
    if (comp) *comp = s->t_comp;
    if (tran) *tran = s->t_tran;
    return 0;
}
long long sarasmg_kernel_launches(sarasmg_handle s) { return s ? s->launches : 0; }
int sarasmg_reset_counters(sarasmg_handle s) {
    ENTER(s);
    TRY(prof_fold(s));
    s->launches = 0;
    memset(s->prof_ms, 0, sizeof(s->prof_ms));
    memset(s->prof_n, 0, sizeof(s->prof_n));
    s->t_comp = s->t_tran = 0.0;
    return 0;
}
int sarasmg_profile_query(sarasmg_handle s, int kind, int level, double *total_ms, long long *launches) {
    ENTER(s);
    if (kind < 0 || kind >= SARASMG_K_COUNT || level < 0 || level > s->D || level >= 32) return set_err("bad kind/level");
    TRY(prof_fold(s));
    if (total_ms) *total_ms = s->prof_ms[kind][level];
    if (launches) *launches = s->prof_n[kind][level];
    return 0;
}
int sarasmg_last_solve_ms(sarasmg_handle s, double *ms) {
    if (!s || !ms) return set_err("null argument");
    *ms = s->last_solve_ms;
    return 0;
}
int sarasmg_set_profile_max_level(sarasmg_handle s, int max_level) {
    if (!s) return set_err("null handle");
    s->prof_max_level = max_level;
    return 0;
}
int sarasmg_event_record(sarasmg_handle s, int slot) {
    ENTER(s);
    if (slot < 0 || slot >= 8) return set_err("sarasmg_event_record: slot must be 0..7");
    if (!s->user_ev[slot]) CU(cudaEventCreate(&s->user_ev[slot]));
    CU(cudaEventRecord(s->user_ev[slot], s->st));
    return 0;
}
int sarasmg_event_elapsed_ms(sarasmg_handle s, int a, int b, double *ms) {
    ENTER(s);
    if (a < 0 || a >= 8 || b < 0 || b >= 8 || !ms || !s->user_ev[a] || !s->user_ev[b]) return set_err("sarasmg_event_elapsed_ms: bad slots");
    CU(cudaEventSynchronize(s->user_ev[b]));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, s->user_ev[a], s->user_ev[b]));
    *ms = f;
    return 0;
}
int sarasmg_debug_set_rhs_mean(sarasmg_handle s, int enable, double value) {
    if (!s) return set_err("null handle");
    s->mean_override = enable ? 1 : 0;
    s->mean_value = value;
    return 0;
}
int sarasmg_last_rhs_mean(sarasmg_handle s, double *value) {
    if (!s || !value) return set_err("sarasmg_last_rhs_mean: null argument");
    *value = s->last_mean;
    return 0;
}
int sarasmg_set_vc_count(sarasmg_handle s, int count) {
    if (!s || count < 0) return set_err("sarasmg_set_vc_count: bad argument");
    s->d.vc_count = count;
    return 0;
}
int sarasmg_set_smoother(sarasmg_handle s, int gs_smooth) {
    if (!s) return set_err("sarasmg_set_smoother: null handle");
    s->d.gs_smooth = gs_smooth ? 1 : 0;
    return 0;
}
int sarasmg_synchronize(sarasmg_handle s) { ENTER(s); CU(cudaStreamSynchronize(s->st)); return 0; }

}  /* extern "C" */
