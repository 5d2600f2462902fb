/*
 * gs_skew.cuh -- exact lexicographic Gauss-Seidel sweep (SARAS lib/poisson/poisson_d3.cc:137-150, poisson_d2.cc:119-128)
 * as ONE kernel launch per sweep.
 *
 * The reference updates p(i,j,k) in place in the order i -> j -> k (k fastest), so a point reads the NEW values of
 * (i-1,j,k), (i,j-1,k), (i,j,k-1) and the OLD values of the three +1 neighbours.  Every execution order that respects
 * that partial order produces bit-identical results.  This kernel uses a skewed pipeline:
 *
 *   - a thread owns one z column (i,j) and walks it in quads of 4 consecutive k (32-byte aligned: one 256-bit load of
 *     p, one of rhs and one 256-bit store per quad; 24 B of HBM traffic per point).  The k recurrence -- the only part
 *     of the dependency chain that cannot be parallelised -- stays in a register of that thread;
 *   - a CTA owns a tile of TI x 32 columns (warp = 32 consecutive columns of the "lane" axis, TI warps along the "warp"
 *     axis; 3-D: lane axis = y, warp axis = x; 2-D: lane axis = x, TI = 1).  At step t thread (a,b) works on quad
 *     t-a-b, i.e. one quad behind its two "minus" neighbours and one ahead of its "plus" neighbours, so that the NEW
 *     quads it needs were produced in the previous step and the OLD quads it needs have not been touched yet.
 *     Neighbour quads travel through a double-buffered shared-memory exchange, one named barrier per step;
 *   - two helper warps per CTA: the halo warp fetches the tile's four face halos (pad planes / neighbour tiles' columns)
 *     two steps ahead with L2-coherent loads and polls the progress counters of the predecessor tiles; the publisher
 *     warp fences and publishes this tile's own progress.  The compute warps never wait on global-memory latency,
 *     flags or fences;
 *   - z is cut into slabs (tasks (I,J,Z)); tasks are handed out through an atomic ticket in an order compatible with the
 *     dependencies (a task only ever waits for tasks with smaller tickets, which are resident or finished: no
 *     deadlock, no grid-wide barrier).  Neighbouring tiles of one slab run concurrently, a few steps apart.
 *
 * Division: the stencil diagonal 4*((cx2+cy2)+cz2) is column-invariant when z is uniform (div_inv with a reciprocal
 * formed once per thread); otherwise it is an IEEE division per point.  Arithmetic and association are those of
 * relax3 / relax2 in mg_kernels.cuh (compiled with -fmad=false).
 */
#pragma once
#include "mg_kernels.cuh"

struct GsPlan {
    int NI, NJ, NZ;         /* tiles along the warp axis, the lane axis, and z slabs */
    int TI;                 /* columns per tile along the warp axis (= compute warps per CTA) */
    int ZQ;                 /* quads per z slab; the last slab also takes the remainder */
    int NQ;                 /* quads per column = ceil(nz / 4) */
    int ntasks;
    int z_uniform;
    int pub;                /* progress is published every `pub` steps (and at the end) */
    const int4 *tasks;      /* (I, J, Z, -) in ticket order */
    int *sync;              /* [0] ticket counter, [1 + (Z*NJ + J)*NI + I] completed steps of task (I,J,Z) */
};

#define GS_TJ 32
#define GS_ROW 34           /* exchange row: 32 lanes + 2 virtual halo lanes */
#define GS_T0 (-3)          /* pipeline warm-up: the first own quad is requested three steps before it is used */

struct __align__(32) quad { double v[4]; };

__device__ __forceinline__ quad ldq_cg(const double *p) {          /* L2-coherent 256-bit load (data written by other SMs in this kernel) */
    quad q;
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q.v[0]), "=d"(q.v[1]), "=d"(q.v[2]), "=d"(q.v[3]) : "l"(p) : "memory");
    return q;
}
__device__ __forceinline__ quad ldq_nc(const double *p) {          /* read-only data (rhs) */
    quad q;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q.v[0]), "=d"(q.v[1]), "=d"(q.v[2]), "=d"(q.v[3]) : "l"(p));
    return q;
}
__device__ __forceinline__ void stq(double *p, const quad &q) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(q.v[0]), "d"(q.v[1]), "d"(q.v[2]), "d"(q.v[3]) : "memory");
}
__device__ __forceinline__ double ld_cg(const double *p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p) {            /* polling load: no L1 invalidation per iteration */
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
/* spin until *p >= need, then acquire */
__device__ __forceinline__ int gs_wait(const int *p, const int need, int seen) {
    if (seen >= need) return seen;
    while ((seen = ld_relaxed(p)) < need) __nanosleep(20);
    return seen;                                        /* the halo data itself is read with L2-coherent (.cg) loads issued after this */
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void gs_bar(const int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__device__ __forceinline__ int gs_task_steps(const GsPlan &G, const int extA, const int extB, const int I, const int J, const int Z) {
    const int ta = min(G.TI, extA - I * G.TI), tb = min(GS_TJ, extB - J * GS_TJ);
    const int nq = (Z == G.NZ - 1) ? G.NQ - Z * G.ZQ : G.ZQ;
    return nq + ta + tb - 2;
}

/* exchange slot: kind 0 = NEW quad a thread produced in the previous step, kind 1 = OLD quad two ahead of it */
#define GS_SLOT(kind, buf, row, e, col) ((((((kind) * 2 + (buf)) * (TI + 2) + (row)) * 4 + (e)) * GS_ROW) + (col))

template <int DIM>
__global__ void __launch_bounds__(256, 2) k_gs_skew(double *P, const double *__restrict__ R, const LevelView L, const GsPlan G) {
    extern __shared__ double gs_sm[];
    __shared__ int s_task;
    __shared__ int s_done;                              /* steps completed by this CTA (halo warp -> publisher warp) */
    const int TI = G.TI;
    const int b = threadIdx.x, a = threadIdx.y;
    const int nthreads = (TI + 1) * 32;                 /* compute warps + halo warp take part in the per-step barrier */
    if (a == TI && b == 0) { s_task = atomicAdd(&G.sync[0], 1); s_done = 0; }
    __syncthreads();
    const int4 tk = G.tasks[s_task];
    const int I = tk.x, J = tk.y, Z = tk.z;
    const int extA = (DIM == 3) ? L.nx : 1, extB = (DIM == 3) ? L.ny : L.nx;
    const int A0 = I * TI, B0 = J * GS_TJ;
    const int ta = min(TI, extA - A0), tb = min(GS_TJ, extB - B0);
    const int q0 = Z * G.ZQ;
    const bool last_slab = (Z == G.NZ - 1);
    const int nq = last_slab ? G.NQ - q0 : G.ZQ;
    const int nsteps = nq + ta + tb - 2;
    const long long strideA = L.plane;                                  /* warp axis (3-D: x) */
    const long long strideB = (DIM == 3) ? (long long)L.pitch : L.plane; /* lane axis (3-D: y, 2-D: x) */
    /* column (A,B) -> offset of its k = -4 element (so that quad q starts at +4q+4... see QOFF) */
    const long long colbase = (DIM == 3) ? L.plane + L.pitch : L.plane + L.pitch;   /* index (0,0): one pad plane + one pad row */
#define GS_COL(A, B) (colbase + (long long)(A) * strideA * (DIM == 3 ? 1 : 0) + (long long)(B) * strideB)
#define GS_QOFF(q) (4LL * (q) + 1 + MG_LEAD)      /* element offset of k = 4q inside a column */
    double *sm_tab = gs_sm + 4 * (TI + 2) * 4 * GS_ROW;                 /* z tables of this slab: cz2[4*nq], cz1[4*nq] */
    const int ntab = 4 * nq;
    {   /* cz2 = c*ztz2[k], cz1 = c*ztzz[k] with c = hxhy (3-D, d3:145-146) or hx2 (2-D, d2:124-125): the leading products of the z terms */
        const double c = (DIM == 3) ? L.hxhy : L.hx2;
        for (int q = a * 32 + b; q < ntab; q += nthreads) {
            const int k = 4 * q0 + q;
            const bool ok = k < L.nz;
            sm_tab[q] = ok ? c * L.ztz2[k] : 0.0;
            sm_tab[2 * G.ZQ * 4 + q] = ok ? c * L.ztzz[k] : 0.0;
        }
    }
    const int idW = (I > 0) ? 1 + (Z * G.NJ + J) * G.NI + (I - 1) : -1;
    const int idL = (J > 0) ? 1 + (Z * G.NJ + (J - 1)) * G.NI + I : -1;
    const int idZ = (Z > 0) ? 1 + ((Z - 1) * G.NJ + J) * G.NI + I : -1;
    int *myprog = G.sync + 1 + (Z * G.NJ + J) * G.NI + I;

    if (a == TI + 1) {
        /* ------------------------------------------------------------------ publisher warp ------------------------- */
        /* Makes this tile's progress visible to the other SMs.  The device-scope fence is slow when the SM has many stores
         * in flight, so it lives in a warp of its own: the step barrier never waits for it, and progress is published as
         * often as the fence allows, always with the latest step count. */
        if (b != 0) return;
        int last = 0;
        while (last < nsteps) {
            int v;
            while ((v = *(volatile int *)&s_done) <= last) __nanosleep(40);
            __threadfence();
            st_release(myprog, v);
            last = v;
        }
        return;
    }
    if (a == TI) {
        /* ------------------------------------------------------------------ halo warp ------------------------------ */
        const int stepsW = (I > 0) ? gs_task_steps(G, extA, extB, I - 1, J, Z) : 0;
        const int stepsL = (J > 0) ? gs_task_steps(G, extA, extB, I, J - 1, Z) : 0;
        if (b == 0 && idZ >= 0) {
            const int stepsZ = gs_task_steps(G, extA, extB, I, J, Z - 1);
            while (ld_acquire(G.sync + idZ) < stepsZ) __nanosleep(100);
        }
        /* this lane's four halo columns */
        const bool doW = (DIM == 3) && b < tb;          /* columns (A0-1, B0+b) [NEW] and (A0+ta, B0+b) [OLD] */
        const bool doL = b < ta;                        /* columns (A0+b, B0-1) [NEW] and (A0+b, B0+tb) [OLD] */
        const double *cWm = P + GS_COL(A0 - 1, B0 + b), *cWp = P + GS_COL(A0 + ta, B0 + b);
        const double *cLm = P + GS_COL(A0 + b, B0 - 1), *cLp = P + GS_COL(A0 + b, B0 + tb);
        int seenW = 0, seenL = 0;
        quad hq[2][4];                                  /* [ring slot][Wm, Wp, Lm, Lp]: loaded in one iteration, stored to the exchange in the next */
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int e = 0; e < 4; e++) hq[r][w].v[e] = 0.0;
        gs_bar(nthreads);                               /* tables written, Z predecessor complete */
        auto halo_step = [&](const int t, quad(&ld)[4], const quad(&st)[4]) {
            /* (1) halos of consumer step c2 = t+2: the producers must have completed their step c2 + lag - 1 */
            const int c2 = t + 2;
            if (b == 0) {
                if (idW >= 0) seenW = gs_wait(G.sync + idW, min(c2 + TI, stepsW), seenW);
                if (idL >= 0) seenL = gs_wait(G.sync + idL, min(c2 + GS_TJ, stepsL), seenL);
            }
            __syncwarp();
            {
                const int qm = c2 - b, qp = c2 - (ta - 1) - b;
                if (doW && qm >= 0 && qm < nq) ld[0] = ldq_cg(cWm + GS_QOFF(q0 + qm));
                if (doW && qp >= 0 && qp < nq) ld[1] = ldq_cg(cWp + GS_QOFF(q0 + qp));
                const int rm = c2 - b, rp = c2 - b - (tb - 1);
                if (doL && rm >= 0 && rm < nq) ld[2] = ldq_cg(cLm + GS_QOFF(q0 + rm));
                if (doL && rp >= 0 && rp < nq) ld[3] = ldq_cg(cLp + GS_QOFF(q0 + rp));
            }
            /* (2) halos of consumer step c1 = t+1 (loaded one iteration ago) into exchange buffer (t+1)&1 */
            {
                const int nb = (t + 1) & 1, c1 = t + 1;
                const int qm = c1 - b, qp = c1 - (ta - 1) - b;
                if (doW && qm >= 0 && qm < nq) {
#pragma unroll
                    for (int e = 0; e < 4; e++) gs_sm[GS_SLOT(0, nb, 0, e, b + 1)] = st[0].v[e];
                }
                if (doW && qp >= 0 && qp < nq) {
#pragma unroll
                    for (int e = 0; e < 4; e++) gs_sm[GS_SLOT(1, nb, ta + 1, e, b + 1)] = st[1].v[e];
                }
                const int rm = c1 - b, rp = c1 - b - (tb - 1);
                if (doL && rm >= 0 && rm < nq) {
#pragma unroll
                    for (int e = 0; e < 4; e++) gs_sm[GS_SLOT(0, nb, b + 1, e, 0)] = st[2].v[e];
                }
                if (doL && rp >= 0 && rp < nq) {
#pragma unroll
                    for (int e = 0; e < 4; e++) gs_sm[GS_SLOT(1, nb, b + 1, e, tb + 1)] = st[3].v[e];
                }
            }
            gs_bar(nthreads);                           /* end of step t: exchange buffer complete, global stores of step t issued */
            if (b == 0 && t >= 0) { __threadfence_block(); *(volatile int *)&s_done = t + 1; }
        };
        for (int t = GS_T0; t < nsteps; t += 2) {       /* unrolled by the ring depth: no register moves behind in-flight loads */
            halo_step(t, hq[0], hq[1]);
            if (t + 1 < nsteps) halo_step(t + 1, hq[1], hq[0]);
        }
        return;
    }

    /* ---------------------------------------------------------------------- compute warps ----------------------- */
    const int Ai = A0 + a, Bi = B0 + b;
    const bool col_ok = (a < ta) && (b < tb);
    const int ix = (DIM == 3) ? Ai : Bi;                /* x index */
    const int jy = (DIM == 3) ? Bi : 0;
    double cx2 = 0.0, cx1 = 0.0, cy2 = 0.0, cy1 = 0.0, two_h, s_xy = 0.0;
    const double hx = L.hx, hy = L.hy, hz = L.hz;
    if (DIM == 3) {
        if (col_ok) { cx2 = L.hyhz * L.xix2[ix]; cx1 = L.hyhz * L.xixx[ix]; cy2 = L.hzhx * L.ety2[jy]; cy1 = L.hzhx * L.etyy[jy]; }
        two_h = 2.0 * L.hxhyhz;
        s_xy = cx2 + cy2;
    } else {
        if (col_ok) { cx2 = L.hz2 * L.xix2[ix]; cx1 = L.hz2 * L.xixx[ix]; }
        two_h = 2.0 * L.hzhx;
        s_xy = cx2;
    }
    const bool zuni = G.z_uniform != 0;
    double den_c = 1.0, rden_c = 1.0;
    const long long mycol = GS_COL(Ai, Bi);
    double *pc = P + mycol;
    const double *rc = R + mycol;
    gs_bar(nthreads);                                   /* tables written, Z predecessor complete */
    if (zuni && col_ok) { den_c = 4.0 * (s_xy + sm_tab[0]); rden_c = 1.0 / den_c; }
    double pB = 0.0;
    if (col_ok) pB = ld_cg(pc + GS_QOFF(q0) - 1);       /* new p(i,j,4*q0-1) of the previous slab, or the k = -1 pad */
    /* highest slab-local quad that may be loaded from the own column: the first quad of the next slab, or (only when nz is a
     * multiple of 4) the quad holding the k = nz pad */
    const int nq_ld = (q0 + nq < G.NQ || (L.nz & 3) == 0) ? nq : nq - 1;
    quad pr[4], rr[2];                                  /* register rings: own OLD quads ql..ql+3 and rhs quads ql, ql+1 */
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int e = 0; e < 4; e++) pr[r].v[e] = 0.0;
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int e = 0; e < 4; e++) rr[r].v[e] = 0.0;

    /* one step; oq/on/onn/on3 = ring slots holding own OLD quads ql, ql+1, ql+2, ql+3 (on3 is requested here, onn was
     * requested one step ago and is published at the end of this step), rq/rqn = rhs quads ql, ql+1 */
    auto step = [&](const int t, const quad &oq, const quad &on, const quad &onn, quad &on3, const quad &rq, quad &rqn) {
        const int ql = t - a - b;
        const int buf = t & 1, nb = buf ^ 1;
        if (col_ok && (ql + 3 >= 0) && (ql + 3 <= nq_ld)) on3 = ldq_cg(pc + GS_QOFF(q0 + ql + 3));
        if (col_ok && (ql + 1 >= 0) && (ql + 1 < nq)) rqn = ldq_nc(rc + GS_QOFF(q0 + ql + 1));
        const bool active = col_ok && ql >= 0 && ql < nq;
        if (active) {
            const int k0 = 4 * (q0 + ql);
            const int nvalid = min(4, L.nz - k0);
            double wm[4], wp[4], lm[4], lp[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                lm[e] = gs_sm[GS_SLOT(0, buf, a + 1, e, b)];
                lp[e] = gs_sm[GS_SLOT(1, buf, a + 1, e, b + 2)];
                if (DIM == 3) {
                    wm[e] = gs_sm[GS_SLOT(0, buf, a, e, b + 1)];
                    wp[e] = gs_sm[GS_SLOT(1, buf, a + 2, e, b + 1)];
                }
            }
            /* the part of the numerator that does not depend on the k recurrence, in the reference's association */
            double txy[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (DIM == 3) {
                    double u = cx2 * (wp[e] + wm[e]) * 2.0;
                    u = u + cx1 * (wp[e] - wm[e]) * hx;
                    u = u + cy2 * (lp[e] + lm[e]) * 2.0;
                    u = u + cy1 * (lp[e] - lm[e]) * hy;
                    txy[e] = u;
                } else {
                    double u = cx2 * (lp[e] + lm[e]) * 2.0;
                    u = u + cx1 * (lp[e] - lm[e]) * hx;
                    txy[e] = u;
                }
            }
            quad o;
            double pb = pB;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const double pT = (e < 3) ? oq.v[e + 1] : on.v[0];
                const double cz2 = sm_tab[4 * ql + e], cz1 = sm_tab[2 * G.ZQ * 4 + 4 * ql + e];
                double u = txy[e] + cz2 * (pT + pb) * 2.0;
                u = u + cz1 * (pT - pb) * hz;
                u = u - two_h * rq.v[e];
                double r;
                if (zuni) r = div_inv(u, den_c, rden_c);
                else r = u / (4.0 * (s_xy + cz2));
                if (e < nvalid) pb = r;
                o.v[e] = r;
            }
            pB = pb;
            if (nvalid == 4) stq(pc + GS_QOFF(q0 + ql), o);
            else
                for (int e = 0; e < nvalid; e++) pc[GS_QOFF(q0 + ql) + e] = o.v[e];
#pragma unroll
            for (int e = 0; e < 4; e++) gs_sm[GS_SLOT(0, nb, a + 1, e, b + 1)] = o.v[e];
        }
        if (col_ok && ql + 2 >= 0 && ql + 2 < nq) {     /* OLD quad ql+2 for the minus neighbours' next step */
#pragma unroll
            for (int e = 0; e < 4; e++) gs_sm[GS_SLOT(1, nb, a + 1, e, b + 1)] = onn.v[e];
        }
        gs_bar(nthreads);
    };
    /* unrolled by the ring depth so that no register is moved behind an in-flight load */
    for (int t = GS_T0; t < nsteps; t += 4) {
        step(t, pr[0], pr[1], pr[2], pr[3], rr[0], rr[1]);
        if (t + 1 < nsteps) step(t + 1, pr[1], pr[2], pr[3], pr[0], rr[1], rr[0]);
        if (t + 2 < nsteps) step(t + 2, pr[2], pr[3], pr[0], pr[1], rr[0], rr[1]);
        if (t + 3 < nsteps) step(t + 3, pr[3], pr[0], pr[1], pr[2], rr[1], rr[0]);
    }
#undef GS_COL
#undef GS_QOFF
}
#undef GS_SLOT
