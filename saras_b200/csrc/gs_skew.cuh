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
    int tabn;               /* doubles per z table = 4 * (quads of the largest slab) */
    int slack;              /* when the halo warp has to wait for a predecessor tile it waits until that tile is `slack` steps further
                               than strictly needed, so that afterwards its polls succeed from the cached progress value instead of
                               meeting every publish of the predecessor (0 = wait for exactly what is needed) */
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
    fence_acq_rel();                                    /* acquire: orders the halo loads issued after this (by this lane and, after __syncwarp, by
                                                           the other lanes) behind the producer's st.release of the progress word */
    return seen;
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

/* ---- shared-memory rings (doubles).  A column slot holds one quad (4 doubles, 32 B, contiguous so that cp.async can fill it).
 *   PO[GS_DP][TI+2][GS_ROW][4]   OLD p quads of the tile's columns and of its plus-side halo columns (row ta+1, col tb+1)
 *   PN[GS_DP][TI+2][GS_ROW][4]   NEW p quads: written by the compute threads, minus-side halo columns (row 0, col 0) by the halo warp
 *   RR[GS_DP][TI][32][4]         rhs quads of the tile's columns
 * quad q of a column lives in ring slot q & (GS_DP-1).                                                                      */
#define GS_DP 4             /* ring depth = prefetch distance of the tile's own columns, in steps */
#define GS_DH 3             /* prefetch distance of the halo columns */
/* a quad is stored as two 16-byte halves, the halves of one row GS_ROW (or 32) columns apart, so that a warp's 128-bit
 * shared-memory accesses touch consecutive 16-byte words (no bank conflicts) and cp.async can fill each half */
#define GS_PO(slot, row, col) (po + ((((slot) * (TI + 2) + (row)) * 2 * GS_ROW + (col)) << 1))
#define GS_PN(slot, row, col) (pn + ((((slot) * (TI + 2) + (row)) * 2 * GS_ROW + (col)) << 1))
#define GS_RR(slot, row, col) (rr + ((((slot) * TI + (row)) * 2 * 32 + (col)) << 1))
#define GS_HP (2 * GS_ROW)  /* doubles between the two halves of a PO / PN quad */
#define GS_HR (2 * 32)      /* ... of an RR quad */

__device__ __forceinline__ void cp_async_quad(double *smem_dst, const int half, const double *gsrc) {   /* 32 B global -> shared, L2 only, asynchronous */
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.cg.shared.global [%2], [%3], 16;"
                 ::"r"(d), "l"(gsrc), "r"(d + 8u * half), "l"(gsrc + 2) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ quad lds_quad(const double *p, const int half) {
    quad q;
    const double2 lo = *reinterpret_cast<const double2 *>(p), hi = *reinterpret_cast<const double2 *>(p + half);
    q.v[0] = lo.x; q.v[1] = lo.y; q.v[2] = hi.x; q.v[3] = hi.y;
    return q;
}
__device__ __forceinline__ void sts_quad(double *p, const int half, const quad &q) {
    *reinterpret_cast<double2 *>(p) = make_double2(q.v[0], q.v[1]);
    *reinterpret_cast<double2 *>(p + half) = make_double2(q.v[2], q.v[3]);
}

template <int DIM>
__global__ void __launch_bounds__(256, 2) k_gs_skew(double *P, const double *__restrict__ R, const LevelView L, const GsPlan G) {
    extern __shared__ __align__(32) double gs_sm[];
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
    const long long colbase = L.plane + L.pitch;                        /* column (0,0): one pad plane + one pad row */
#define GS_COL(A, B) (colbase + (long long)(A) * strideA * (DIM == 3 ? 1 : 0) + (long long)(B) * strideB)
#define GS_QOFF(q) (4LL * (q) + 1 + MG_LEAD)      /* element offset of k = 4q inside a column */
    double *po = gs_sm;
    double *pn = po + GS_DP * (TI + 2) * GS_ROW * 4;
    double *rr = pn + GS_DP * (TI + 2) * GS_ROW * 4;
    double *sm_tab = rr + GS_DP * TI * 32 * 4;                          /* z tables of this slab: cz2[4*nq], cz1[4*nq] */
    const int ntab = 4 * nq;
    {   /* cz2 = c*ztz2[k], cz1 = c*ztzz[k] with c = hxhy (3-D, d3:145-146) or hx2 (2-D, d2:124-125): the leading products of the z terms */
        const double c = (DIM == 3) ? L.hxhy : L.hx2;
        for (int q = a * 32 + b; q < ntab; q += (TI + 2) * 32) {
            const int k = 4 * q0 + q;
            const bool ok = k < L.nz;
            sm_tab[q] = ok ? c * L.ztz2[k] : 0.0;
            sm_tab[G.tabn + q] = ok ? c * L.ztzz[k] : 0.0;
        }
    }
    const int idW = (I > 0) ? 1 + (Z * G.NJ + J) * G.NI + (I - 1) : -1;
    const int idL = (J > 0) ? 1 + (Z * G.NJ + (J - 1)) * G.NI + I : -1;
    const int idZ = (Z > 0) ? 1 + ((Z - 1) * G.NJ + J) * G.NI + I : -1;
    int *myprog = G.sync + 1 + (Z * G.NJ + J) * G.NI + I;
    /* highest slab-local quad that may be loaded from a column: the first quad of the next slab, or (only when nz is a
     * multiple of 4) the quad holding the k = nz pad */
    const int nq_ld = (q0 + nq < G.NQ || (L.nz & 3) == 0) ? nq : nq - 1;
    __syncthreads();                                    /* tables complete */

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
        gs_bar(nthreads);                               /* Z predecessor complete */
        for (int t = -GS_DP; t < nsteps; ++t) {
            /* halos of consumer step c = t + GS_DH: the producers must have completed their step c + lag - 1 */
            const int c = t + GS_DH;
            if (b == 0) {
                if (idW >= 0 && seenW < min(c + TI, stepsW)) seenW = gs_wait(G.sync + idW, min(c + TI + G.slack, stepsW), seenW);
                if (idL >= 0 && seenL < min(c + GS_TJ, stepsL)) seenL = gs_wait(G.sync + idL, min(c + GS_TJ + G.slack, stepsL), seenL);
            }
            __syncwarp();
            {
                const int qm = c - b, qp = c - (ta - 1) - b;        /* quads of consumers (0,b) and (ta-1,b) at step c */
                if (doW && qm >= 0 && qm < nq) cp_async_quad(GS_PN(qm & (GS_DP - 1), 0, b + 1), GS_HP, cWm + GS_QOFF(q0 + qm));
                if (doW && qp >= 0 && qp < nq) cp_async_quad(GS_PO(qp & (GS_DP - 1), ta + 1, b + 1), GS_HP, cWp + GS_QOFF(q0 + qp));
                const int rm = c - b, rp = c - b - (tb - 1);        /* lane b serves row b: consumers (b,0) and (b,tb-1) */
                if (doL && rm >= 0 && rm < nq) cp_async_quad(GS_PN(rm & (GS_DP - 1), b + 1, 0), GS_HP, cLm + GS_QOFF(q0 + rm));
                if (doL && rp >= 0 && rp < nq) cp_async_quad(GS_PO(rp & (GS_DP - 1), b + 1, tb + 1), GS_HP, cLp + GS_QOFF(q0 + rp));
            }
            cp_async_commit();
            cp_async_wait<GS_DH - 1>();                 /* halos of consumer step t+1 have landed */
            gs_bar(nthreads);                           /* end of step t: rings complete for step t+1, global stores of step t issued */
            if (b == 0 && t >= 0) { __threadfence_block(); *(volatile int *)&s_done = t + 1; }
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
    if (zuni && col_ok) { den_c = 4.0 * (s_xy + sm_tab[0]); rden_c = 1.0 / den_c; }
    const long long mycol = GS_COL(Ai, Bi);
    double *pc = P + mycol;
    const double *rc = R + mycol;
    gs_bar(nthreads);                                   /* Z predecessor complete */
    double pB = 0.0;
    if (col_ok) pB = ld_cg(pc + GS_QOFF(q0) - 1);       /* new p(i,j,4*q0-1) of the previous slab, or the k = -1 pad */
    /* shared-memory offsets (doubles) of this thread's column inside one ring slot, and ring strides */
    const int row_p = 2 * GS_ROW * 2, slot_p = (TI + 2) * row_p, own_p = (a + 1) * row_p + (b + 1) * 2;
    const int slot_r = TI * 2 * 32 * 2, own_r = a * 2 * 32 * 2 + b * 2;
    quad cur;                                           /* own OLD quad ql (read from the ring one step earlier as quad ql+1) */
#pragma unroll
    for (int e = 0; e < 4; e++) cur.v[e] = 0.0;

    for (int t = -GS_DP; t < nsteps; ++t) {
        const int ql = t - a - b;
        const bool active = col_ok && ql >= 0 && ql < nq;
        if (active) {
            const int k0 = 4 * (q0 + ql);
            const int nvalid = min(4, L.nz - k0);
            const int sl = ql & (GS_DP - 1), sl1 = (ql + 1) & (GS_DP - 1);
            const int so = sl * slot_p + own_p, so1 = sl1 * slot_p + own_p;      /* own column in ring slots of quads ql, ql+1 */
            if (ql == 0) cur = lds_quad(po + so, GS_HP);
            const quad nxt = lds_quad(po + so1, GS_HP);                  /* own OLD quad ql+1 (.x = pT of the last point) */
            const quad lm = lds_quad(pn + so - 2, GS_HP);                /* NEW, lane-minus neighbour */
            const quad lp = lds_quad(po + so + 2, GS_HP);                /* OLD, lane-plus neighbour */
            const quad rq = lds_quad(rr + sl * slot_r + own_r, GS_HR);
            const quad tz2 = lds_quad(sm_tab + 4 * ql, 2), tz1 = lds_quad(sm_tab + G.tabn + 4 * ql, 2);
            /* the part of the numerator that does not depend on the k recurrence, in the reference's association */
            double txy[4];
            if (DIM == 3) {
                const quad wm = lds_quad(pn + so - row_p, GS_HP);        /* NEW, warp-minus neighbour */
                const quad wp = lds_quad(po + so + row_p, GS_HP);        /* OLD, warp-plus neighbour */
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    double u = cx2 * (wp.v[e] + wm.v[e]) * 2.0;
                    u = u + cx1 * (wp.v[e] - wm.v[e]) * hx;
                    u = u + cy2 * (lp.v[e] + lm.v[e]) * 2.0;
                    u = u + cy1 * (lp.v[e] - lm.v[e]) * hy;
                    txy[e] = u;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    double u = cx2 * (lp.v[e] + lm.v[e]) * 2.0;
                    u = u + cx1 * (lp.v[e] - lm.v[e]) * hx;
                    txy[e] = u;
                }
            }
            /* stencil diagonal and its correctly rounded reciprocal (off the recurrence) */
            double den[4], rden[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (zuni) { den[e] = den_c; rden[e] = rden_c; }
                else { den[e] = 4.0 * (s_xy + tz2.v[e]); rden[e] = 1.0 / den[e]; }
            }
            /* the k recurrence: straight-line code; operands outside div_inv's exact range only raise a flag */
            quad o;
            double pb = pB;
            unsigned bad = 0u;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const double pT = (e < 3) ? cur.v[e + 1] : nxt.v[0];
                double u = txy[e] + tz2.v[e] * (pT + pb) * 2.0;
                u = u + tz1.v[e] * (pT - pb) * hz;
                u = u - two_h * rq.v[e];
                const double r = div_inv_fast(u, den[e], rden[e]);
                bad |= div_inv_bad(u);
                if (e < nvalid) pb = r;
                o.v[e] = r;
            }
            if (bad) {                                   /* rare: redo the quad with IEEE divisions */
                pb = pB;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const double pT = (e < 3) ? cur.v[e + 1] : nxt.v[0];
                    double u = txy[e] + tz2.v[e] * (pT + pb) * 2.0;
                    u = u + tz1.v[e] * (pT - pb) * hz;
                    u = u - two_h * rq.v[e];
                    const double r = div_ieee(u, den[e]);
                    if (e < nvalid) pb = r;
                    o.v[e] = r;
                }
            }
            pB = pb;
            if (nvalid == 4) stq(pc + GS_QOFF(q0 + ql), o);
            else
                for (int e = 0; e < nvalid; e++) pc[GS_QOFF(q0 + ql) + e] = o.v[e];
            sts_quad(pn + so, GS_HP, o);
            cur = nxt;
        }
        /* prefetch GS_DP steps ahead: own OLD quad and rhs quad ql + GS_DP go into the ring slots of quad ql, which nobody
         * reads any more (this thread just did, its minus neighbours did one step ago) */
        {
            const int qn = ql + GS_DP;
            if (col_ok && qn >= 0 && qn <= nq_ld) cp_async_quad(GS_PO(qn & (GS_DP - 1), a + 1, b + 1), GS_HP, pc + GS_QOFF(q0 + qn));
            if (col_ok && qn >= 0 && qn < nq) cp_async_quad(GS_RR(qn & (GS_DP - 1), a, b), GS_HR, rc + GS_QOFF(q0 + qn));
            cp_async_commit();
        }
        cp_async_wait<GS_DP - 2>();                     /* own quad ql+2 (read by the minus neighbours in the next step) has landed */
        gs_bar(nthreads);
    }
#undef GS_COL
#undef GS_QOFF
}
#undef GS_PO
#undef GS_PN
#undef GS_RR
