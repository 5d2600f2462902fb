/*
 * gs_strip.cuh -- exact lexicographic Gauss-Seidel sweep for the large 3-D levels
 * (SARAS lib/poisson/poisson_d3.cc:137-150), one launch per sweep, without any CTA-wide barrier.
 *
 * The reference updates p(i,j,k) in place in the order i -> j -> k (k fastest): a point reads the NEW values of
 * (i-1,j,k), (i,j-1,k), (i,j,k-1) and the OLD values of the three +1 neighbours.  Any execution order that respects that
 * partial order gives the same bits.  k_gs_skew (gs_skew.cuh) advances a whole CTA one quad per named barrier and is
 * paced by that barrier and by one cross-SM hand-off per step; here every WARP is a self-contained pipeline:
 *
 *   - a warp owns a tile of 32 columns in y (a "strip") x nq quads in z (a "slab", 4 nq points) and marches it through
 *     ALL x planes.  Lane b owns column j = 32 J + b and runs one step (= one quad: 256-bit load of p and rhs, 256-bit
 *     store) behind lane b-1: at step t it works on position t - b of the sequence (plane 0 quad 0, plane 0 quad 1, ...,
 *     plane 1 quad 0, ...).  Consecutive planes follow each other without a bubble;
 *   - the warp keeps a WINDOW of its tile in shared memory: ring slot (p mod RW) holds, for every column of the strip
 *     (plus one halo column on either side), the quad of position p -- its OLD value from nq + DP positions before the
 *     owning lane reaches p (cp.async prefetch, L2 only) until that lane updates it in place, its NEW value until one
 *     plane later.  A lane at position p then finds everything in the window: +x = slot(p + nq) (OLD), -x = slot(p - nq)
 *     (NEW), +z = slot(p + 1), and -y / +y in its OWN slot row one cell to the left (lane b-1 wrote its NEW quad of p one
 *     step ago) and to the right (lane b+1 reaches p in the next step: still OLD).  One __syncwarp per step orders
 *     these exchanges; the k recurrence and the -z value stay in registers;
 *   - what crosses warps: (1) the halo columns: lane 0 requests column 32J-1 (NEW, so only after tile (J-1, zw) has
 *     published enough progress), the last live lane column 32J+nb (OLD), one quad per step each, into the halo cells of
 *     the window; (2) the -z value of a slab's first quad (the top point of the slab below: shared memory when that slab
 *     lives in the same CTA, global memory otherwise) and the OLD first point of the slab above.  Both dependencies are
 *     feed-forward (strip J-1 -> J, slab zw-1 -> zw), so latency only adds start-up skew;
 *   - progress: a compute warp writes its completed step count to shared memory; the CTA's comm warp -- which has no
 *     stores of its own, so its device-scope fence is cheap -- publishes the counts to global memory and polls the counts
 *     of the predecessor tiles in other CTAs on behalf of the compute warps.  CTAs take tickets in a
 *     dependency-compatible order, so a tile only ever waits for tiles that are already resident or finished.
 *
 * The step body is kept short on purpose (the sweep is issue-bound before it is HBM-bound): 32-bit shared-space addresses
 * advanced incrementally, running 64-bit global pointers, warp-uniform wait conditions, no shuffles.
 * Arithmetic, association and the exact-division scheme are those of k_gs_skew (relax3 in mg_kernels.cuh, -fmad=false).
 * scripts/sim/gs_strip_emul.py restates this schedule on the CPU and checks it against sequential Gauss-Seidel.
 */
#pragma once
#include "gs_skew.cuh"

struct GsStripPlan {
    int NJ;                 /* strips of 32 columns along y */
    int NZW;                /* slabs along z */
    int NG;                 /* slab groups: a CTA runs WC consecutive slabs of one strip */
    int nq;                 /* quads per slab; the last slab holds the remainder (1..nq) */
    int NQ;                 /* quads per column = ceil(nz / 4) */
    int WC;                 /* compute warps per CTA */
    int DP;                 /* prefetch distance in steps */
    int RW;                 /* window ring: 2 nq + DP + 1 slots */
    int RRn;                /* rhs ring: DP + 1 slots */
    int ntasks;             /* NJ * NG */
    int z_uniform;
    int warp_bytes;
    const double *xpair;    /* {xix2[i], xixx[i]} interleaved, i = 0 .. nx-1 (one 16-byte request per plane) */
    const double *rden;     /* z-uniform levels: RN(1 / (4 ((hyhz xix2[i] + hzhx ety2[j]) + hxhy ztz2[0]))) per column, [nx][ny] */
    double two_h;           /* 2 hxhyhz of the level: operands that never change are read from the parameter bank, not held in registers */         /* shared-memory bytes per compute warp */
    const int2 *tasks;      /* (J, G) in ticket order */
    int *sync;              /* [0] ticket, [1 + zw*NJ + J] completed steps of tile (J, zw) */
};

#define GSS_ZD 2            /* planes of -z hand-off values buffered between two slabs of a CTA (power of two) */
#ifndef GSS_SLEEP
#define GSS_SLEEP(ns) __nanosleep(ns)
#endif
#define GSS_MAXW 13         /* compute warps per CTA: 13 + the comm warp = 448 threads */
#define GSS_CELLS 34        /* cells of a window row: 32 lanes + a halo cell on either side */
#define GSS_HALF (GSS_CELLS * 16)       /* bytes between the two 16-byte halves of a quad (a row of half-quads) */
#define GSS_SLOT (2 * GSS_HALF)         /* bytes per window slot */
#define GSS_RSLOT 1024                  /* bytes per rhs-ring slot: [2 halves][32 lanes] x 16 B */

/* side rings: what a lane needs when it enters a plane (x coefficients, reciprocal diagonal, -z value) or reaches its top
 * quad (the OLD point above the slab), requested DP steps ahead like everything else.  SD planes per lane, XW planes of x
 * coefficients per warp (lane 0 requests them, lane 31 reads them 31 steps later) */
__host__ __device__ __forceinline__ int gss_side_planes(const int nq, const int DP) { return (nq - 1 + DP) / nq + 1; }
__host__ __device__ __forceinline__ int gss_x_planes(const int nq, const int DP) {
    int w = 2;
    while (w < (31 + DP) / nq + 2) w *= 2;
    return w;
}
__host__ __device__ __forceinline__ int gss_warp_bytes(const int nq, const int DP) {
    /* window ring, rhs ring, -z hand-off ring, z tables (cz2, cz1), x coefficients, three per-lane side rings (rden 8 B,
     * pTop 8 B, pB 16 B per lane and plane) */
    return (2 * nq + DP + 1) * GSS_SLOT + (DP + 1) * GSS_RSLOT + GSS_ZD * 32 * 8 + 2 * 4 * nq * 8 + gss_x_planes(nq, DP) * 16 +
           gss_side_planes(nq, DP) * 32 * (8 + 8 + 16);
}

__device__ __forceinline__ void st_relaxed_gpu(int *p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
/* shared-space accesses with 32-bit addresses */
__device__ __forceinline__ quad lds_q(const unsigned a, const unsigned half) {
    quad q;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(q.v[0]), "=d"(q.v[1]) : "r"(a) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(q.v[2]), "=d"(q.v[3]) : "r"(a + half) : "memory");
    return q;
}
__device__ __forceinline__ void sts_q(const unsigned a, const unsigned half, const quad &q) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(q.v[0]), "d"(q.v[1]) : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a + half), "d"(q.v[2]), "d"(q.v[3]) : "memory");
}
__device__ __forceinline__ void cpa_q(const unsigned dst, const unsigned half, const char *gsrc) {      /* 32 B global -> shared, L2 only, asynchronous */
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.cg.shared.global [%2], [%3], 16;"
                 ::"r"(dst), "l"(gsrc), "r"(dst + half), "l"(gsrc + 16) : "memory");
}
__device__ __forceinline__ double ld_cg_b(const char *p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void cpa_8(const unsigned dst, const void *gsrc) {       /* 8 B global -> shared through L1 (static or first-touch data) */
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa_16(const unsigned dst, const void *gsrc) {      /* 16 B global -> shared, L2 only */
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ double lds_f64(const unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds_vf64(const unsigned a) {
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
/* acquire / release at CTA scope (MEMBAR.ALL.CTA; __threadfence_block() is the heavier sequentially consistent MEMBAR.SC.CTA) */
__device__ __forceinline__ void gss_fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
/* spin on a shared-memory progress word.  No __nanosleep here: these waits are a fraction of a step long and recur every
 * step, the load's own latency paces the loop */
__device__ __forceinline__ int gss_spin(const volatile int *w, const int need) {
    int v;
    while ((v = *w) < need) { }
    return v;
}
/* the same for words that may lag far behind (start-up skew of the strips): back off once the gap is more than a few steps */
__device__ __forceinline__ int gss_wait_far(const volatile int *w, const int need) {
    int v;
    while ((v = *w) < need) {
        if (need - v > 6) GSS_SLEEP(500);
    }
    return v;
}

/* the k recurrence of one quad (the x / y part of the numerator, txy, is formed before).
 * PARTIAL: the column's last quad with nvalid < 4 live points (topmost slab only) */
template <bool ZUNI, bool PARTIAL>
__device__ __forceinline__ void gss_chain(const double (&txy)[4], const quad &cur, const double pT3, const quad &rq, const quad &tz2,
                                          const quad &tz1, const double hz, const double two_h, const double s_xy, const double den_c,
                                          const double rden_c, const int nvalid, double &pB, quad &o) {
    double pb = pB;
    unsigned bad = 0u;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        /* the diagonal and its correctly rounded reciprocal do not depend on the recurrence */
        const double den = ZUNI ? den_c : 4.0 * (s_xy + tz2.v[e]);
        const double rden = ZUNI ? rden_c : __drcp_rn(den);
        const double pT = (e < 3) ? cur.v[e + 1] : pT3;
        double u = txy[e] + tz2.v[e] * (pT + pb) * 2.0;
        u = u + tz1.v[e] * (pT - pb) * hz;
        u = u - two_h * rq.v[e];
        const double r = div_inv_fast(u, den, rden);
        bad |= div_inv_bad(u);
        if (!PARTIAL || e < nvalid) pb = r;
        o.v[e] = r;
    }
    if (bad) {                                          /* rare: operands outside div_inv's exact range: redo with IEEE divisions */
        pb = pB;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const double pT = (e < 3) ? cur.v[e + 1] : pT3;
            double u = txy[e] + tz2.v[e] * (pT + pb) * 2.0;
            u = u + tz1.v[e] * (pT - pb) * hz;
            u = u - two_h * rq.v[e];
            const double r = div_ieee(u, ZUNI ? den_c : 4.0 * (s_xy + tz2.v[e]));
            if (!PARTIAL || e < nvalid) pb = r;
            o.v[e] = r;
        }
    }
    pB = pb;
}

/* the reciprocal of the stencil diagonal of every column of a z-uniform level: formed once per level with the same
 * operations the sweep would use (correctly rounded reciprocal), so that the sweep needs no MUFU and no Newton steps */
__global__ void __launch_bounds__(256) k_gss_rden(const LevelView L, double *__restrict__ out) {
    const long long n = (long long)L.nx * L.ny;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / L.ny), j = (int)(t % L.ny);
        const double cx2 = L.hyhz * L.xix2[i], cy2 = L.hzhx * L.ety2[j], cz2 = L.hxhy * L.ztz2[0];
        const double s_xy = cx2 + cy2;
        out[t] = __drcp_rn(4.0 * (s_xy + cz2));
    }
}

__global__ void __launch_bounds__(256) k_gss_xpair(const LevelView L, double *__restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.nx; i += gridDim.x * blockDim.x) {
        out[2 * i] = L.xix2[i];
        out[2 * i + 1] = L.xixx[i];
    }
}

template <bool ZUNI>
__global__ void __launch_bounds__(32 * (GSS_MAXW + 1), 1) k_gs_strip(double *P, const double *__restrict__ R, const LevelView L, const GsStripPlan G) {
    extern __shared__ __align__(128) unsigned char gss_sm[];
    __shared__ int s_task;
    __shared__ volatile int s_done[GSS_MAXW + 1];       /* steps completed by compute warp a */
    __shared__ volatile int s_seenY[GSS_MAXW + 1];      /* published progress of the tile (J-1, zw_a), as last seen by the comm warp */
    __shared__ volatile int s_seenZ0;                   /* published progress of the tile below warp 0's slab (other CTA) */
    const int b = threadIdx.x, a = threadIdx.y;
    const int WC = G.WC;
    if (a == WC) {
        if (b == 0) { s_task = atomicAdd(&G.sync[0], 1); s_seenZ0 = 0; }
        if (b <= GSS_MAXW) { s_done[b] = 0; s_seenY[b] = 0; }
    }
    __syncthreads();                                    /* the only CTA-wide barrier of the kernel */
    const int2 tk = G.tasks[s_task];
    const int J = tk.x, Gz = tk.y;
    const int nx = L.nx, nq = G.nq, NJ = G.NJ;
    const int nb = min(32, L.ny - 32 * J);              /* live lanes of this strip */
    const int Pend = nx * nq;                           /* positions of a lane's sequence: every slab walks nq per plane */

    if (a == WC) {
        /* ------------------------------------------------------------------ comm warp ------------------------------ */
        /* lane w serves compute warp w: publishes its step count, polls the tile (J-1, zw_w); lane 0 also polls the tile
         * below the CTA's first slab.  One device-scope fence per round orders [counts read from shared memory, words polled]
         * before [words published, values handed to the compute warps]. */
        const int zw = Gz * WC + b;
        const bool serve = (b < WC) && (zw < G.NZW);
        const int final_my = Pend + nb - 1;
        const int final_y = Pend + 31;                  /* strips 0 .. NJ-2 are full */
        int *w_my = G.sync + 1 + (serve ? zw : 0) * NJ + J;
        const int *w_y = (serve && J > 0) ? G.sync + 1 + zw * NJ + (J - 1) : nullptr;
        const int *w_z = (b == 0 && Gz > 0) ? G.sync + 1 + (Gz * WC - 1) * NJ + J : nullptr;
        int pub = 0, seen_y = 0, seen_z = 0;
        while (true) {
            const int d = serve ? s_done[b] : 0;
            int y = seen_y, z = seen_z;
            if (w_y && seen_y < final_y) y = ld_relaxed(w_y);
            if (w_z && seen_z < final_my) z = ld_relaxed(w_z);
            gss_fence_cta();
            fence_acq_rel();
            if (serve && d > pub) { st_relaxed_gpu(w_my, d); pub = d; }
            if (y > seen_y) { seen_y = y; s_seenY[b] = y; }
            if (z > seen_z) { seen_z = z; s_seenZ0 = z; }
            const bool fin = !serve || pub >= final_my;
            if (__all_sync(0xffffffffu, fin)) break;
            GSS_SLEEP(100);
        }
        return;
    }

    /* ---------------------------------------------------------------------- compute warps ----------------------- */
    const int zw = Gz * WC + a;
    if (zw >= G.NZW) return;                            /* last group: fewer slabs than warps */
    const int q0 = zw * nq;                             /* first quad of the slab */
    /* live quads of this slab: nq, fewer only in the topmost slab, which idles through its dead quads q >= nqw.  All tiles
     * thus share one step structure: a lane runs between nq + 1 and (GSS_ZD + 1) nq - 1 steps behind the same lane of the
     * slab below, whatever the remainder NQ mod nq is */
    const int nqw = min(nq, G.NQ - q0);
    const int DP = G.DP;
    const bool has_below_cta = (a > 0);                 /* slab below runs in warp a-1 of this CTA */
    const bool has_below = has_below_cta || (Gz > 0);   /* ... or in another CTA */
    const bool has_above_cta = (a + 1 < WC) && (zw + 1 < G.NZW);
    const bool has_y = (J > 0);
    const int final_all = Pend + nb - 1;                /* step count every tile of this strip ends with */

    /* shared-memory map of this warp (32-bit shared-space addresses) */
    const unsigned wbase = (unsigned)__cvta_generic_to_shared(gss_sm) + (unsigned)a * (unsigned)G.warp_bytes;
    const unsigned RWB = (unsigned)G.RW * GSS_SLOT, RRB = (unsigned)G.RRn * GSS_RSLOT;
    const unsigned win_lo = wbase + (unsigned)(b + 1) * 16u;                /* this lane's cell in slot 0, first half */
    const unsigned win_hi = win_lo + RWB;
    const unsigned rr_lo = wbase + RWB + (unsigned)b * 16u, rr_hi = rr_lo + RRB;
    const unsigned zb_mine = wbase + RWB + RRB;                             /* [GSS_ZD][32] doubles handed to THIS warp by warp a-1 */
    const unsigned ztab = zb_mine + GSS_ZD * 256u;                          /* cz2[4 nq], cz1[4 nq] */
    const unsigned zb_above = zb_mine + (unsigned)G.warp_bytes;             /* the hand-off ring of warp a+1 */
    const int SD = gss_side_planes(nq, G.DP), XW = gss_x_planes(nq, G.DP);
    const unsigned xtab = ztab + 64u * (unsigned)nq;                        /* [XW] {xix2, xixx} of plane (i mod XW) */
    const unsigned rd_lo = xtab + 16u * (unsigned)XW + 8u * (unsigned)b;    /* [SD][32] RN(1/diagonal) of the lane's column */
    const unsigned pt_lo = rd_lo + 256u * (unsigned)SD;                     /* [SD][32] OLD first point above the slab */
    const unsigned pb_lo = pt_lo + 256u * (unsigned)SD - 8u * (unsigned)b + 16u * (unsigned)b;      /* [SD][32] 16 B: {k = 4q0-2, k = 4q0-1} below the slab */
    {   /* cz2 = hxhy*ztz2[k], cz1 = hxhy*ztzz[k]: the leading products of the z terms (d3:145-146) */
        for (int e = b; e < 4 * nqw; e += 32) {
            const int k = 4 * q0 + e;
            const bool ok = k < L.nz;
            const double c2 = ok ? L.hxhy * L.ztz2[k] : 0.0, c1 = ok ? L.hxhy * L.ztzz[k] : 0.0;
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(ztab + 8u * e), "d"(c2) : "memory");
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(ztab + 8u * (4 * nq + e)), "d"(c1) : "memory");
        }
    }
    __syncwarp();

    const bool col_ok = b < nb;
    const int j = 32 * J + (col_ok ? b : nb - 1);       /* dead lanes alias the last live column; they never touch memory */
    const long long planeB = L.plane * 8LL, pitchB = (long long)L.pitch * 8LL;
    const long long coloff = L.plane + (long long)(j + 1) * L.pitch + (1 + MG_LEAD) + 4LL * q0;        /* element (0, j, 4 q0) */
    const double cy2 = L.hzhx * L.ety2[j], cy1 = L.hzhx * L.etyy[j];
    double cz2_0, cz1_0;                                /* z-uniform levels: the z coefficients of every live point */
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cz2_0) : "r"(ztab) : "memory");
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cz1_0) : "r"(ztab + 32u * nq) : "memory");
    const int nvalid_last = L.nz - 4 * (G.NQ - 1);      /* live points of the column's last quad (1..4) */
    const int qpart = (zw == G.NZW - 1 && nvalid_last < 4) ? (G.NQ - 1 - q0) : -1;        /* slab-local index of a partial quad */

    /* lane state.  pos = t - b; (i, q) = plane and slab-local quad of pos; (ir, qr) the same for pos + DP */
    const int t0 = -(2 * nq + DP);
    int i, q, ir, qr;
    unsigned aw, ar;                                    /* own cell in the window slot / rhs-ring slot of pos */
    const char *gw;                                     /* P at (plane ir+1, quad qr): the window prefetch */
    char *gs;                                           /* P at (plane i, quad q): the store */
    {
        const int pos = t0 - b;
        const int m = -pos;                             /* > 0 */
        i = -((m + nq - 1) / nq);
        q = pos - i * nq;
        const int pr = pos + DP, mr = -pr;              /* pr < 0 here as well */
        ir = -((mr + nq - 1) / nq);
        qr = pr - ir * nq;
        aw = win_lo + (unsigned)(((pos % G.RW) + G.RW) % G.RW) * GSS_SLOT;
        ar = rr_lo + (unsigned)(((pos % G.RRn) + G.RRn) % G.RRn) * GSS_RSLOT;
        gw = (const char *)(P + coloff) + (long long)(ir + 1) * planeB + 32LL * qr;
        gs = (char *)(P + coloff) + (long long)i * planeB + 32LL * q;
    }
    int sdi, sdr;                                       /* side-ring slot of plane i / plane ir */
    {
        const int pos = t0 - b, pr = pos + DP;
        const int i0 = -((-pos + nq - 1) / nq), ir0 = -((-pr + nq - 1) / nq);
        sdi = ((i0 % SD) + SD) % SD;
        sdr = ((ir0 % SD) + SD) % SD;
    }
    const long long dRP = ((const char *)R - (const char *)P) - planeB;      /* gw -> rhs quad of (plane ir, quad qr) */
    const long long dHalo = -planeB + ((b == 0) ? -pitchB : pitchB);          /* gw -> the lane's halo column, plane ir */
    const unsigned halo_cell = (b == 0) ? (unsigned)-16 : 16u;               /* halo cell relative to the own cell */
    const bool halo_lane = col_ok && (b == 0 || b == nb - 1);
    const long long wrapB = planeB - 32LL * nq;         /* extra pointer advance when a plane ends */
    const unsigned offP = (unsigned)nq * GSS_SLOT, offF = (unsigned)(nq + DP) * GSS_SLOT, offR = (unsigned)DP * GSS_RSLOT;

    double cx2 = 0.0, cx1 = 0.0, s_xy = 0.0, den_c = 1.0, rden_c = 1.0;
    double pTop = 0.0, pB = 0.0;
    int seenY = 0, seenZ = 0, seenA = 0;                /* progress of the -y tile, the -z tile, and my +z consumer (back-pressure) */

    for (int t = t0; t < final_all; ++t) {
        /* ---- (A) dependencies of this step: all warp-uniform ---- */
        {
            /* -y halo: lane 0 requests the NEW quad of position t + DP of column 32J-1, finished by the last lane of tile
             * (J-1, zw) in its step t + DP + 31 */
#ifdef GSS_EXPERIMENT_NO_YWAIT      /* timing experiment only: wrong results */
            const int needY = 0;
#else
            const int needY = (has_y && (unsigned)(t + DP) < (unsigned)Pend) ? t + DP + 32 : 0;
#endif
            if (seenY < needY) {
                seenY = gss_wait_far(&s_seenY[a], needY);
                gss_fence_cta();
            }
            /* -z value of the plane a lane enters in THIS step (pos = i*nq): the slab below finishes plane i in lane b with its
             * step i*nq + nq - 1 + b = t + nq - 1, the same for every lane.  Read from shared memory in this step when that slab
             * lives in this CTA; requested from global memory DP steps ahead otherwise */
#ifdef GSS_EXPERIMENT_NO_ZWAIT      /* timing experiment only: wrong results */
            const int needZ = 0;
#else
            const int needZ = has_below ? min(t + nq + (has_below_cta ? 0 : DP), final_all) : 0;
#endif
            if (seenZ < needZ) {
                seenZ = has_below_cta ? gss_spin(&s_done[a - 1], needZ) : gss_wait_far(&s_seenZ0, needZ);
                gss_fence_cta();
            }
            /* back-pressure on the -z hand-off ring: a lane replaces the value of plane i - GSS_ZD by plane i's at its top quad
             * (step t = i*nq + nq - 1 + b); the slab above must have read it, which its lane b does in step (i - GSS_ZD)*nq + b */
            const int needA = has_above_cta ? max(t + 2 - (GSS_ZD + 1) * nq, 0) : 0;
            if (seenA < needA) {
                seenA = gss_spin(&s_done[a + 1], needA);
                gss_fence_cta();
            }
        }

        /* ---- (B) prefetch: window quad of plane ir + 1 (position pos + nq + DP); rhs and halo quads of position pos + DP ---- */
        {
            const bool qok = col_ok && qr < nqw;
            unsigned af = aw + offF; af = (af >= win_hi) ? af - RWB : af; af = (af >= win_hi) ? af - RWB : af;
            if (qok && (unsigned)(ir + 2) < (unsigned)(nx + 2)) cpa_q(af, GSS_HALF, gw);
            if (qok && (unsigned)ir < (unsigned)nx) {
                unsigned arf = ar + offR; arf = (arf >= rr_hi) ? arf - RRB : arf;
                cpa_q(arf, 512u, gw + dRP);
                /* halo columns of position pos + DP into the halo cell of ITS slot (= the prefetch slot moved back one plane) */
                if (halo_lane) {
                    unsigned ah = af + (RWB - offP); ah = (ah >= win_hi) ? ah - RWB : ah;
                    cpa_q(ah + halo_cell, GSS_HALF, gw + dHalo);
                    if (nb == 1) cpa_q(ah + 16u, GSS_HALF, gw - planeB + pitchB);       /* a one-column strip: lane 0 is the last lane too */
                }
            }
            /* side rings, position pos + DP: entering plane ir / at its top quad */
            if (qok && (unsigned)ir < (unsigned)nx) {
                if (qr == 0) {
                    if (ZUNI) cpa_8(rd_lo + 256u * (unsigned)sdr, G.rden + (long long)ir * L.ny + j);
                    if (!has_below_cta) cpa_16(pb_lo + 512u * (unsigned)sdr, gw - planeB - 16);     /* {k = 4q0-2, 4q0-1}: the z pad (static) or the top of the slab below */
                    if (b == 0) cpa_16(xtab + 16u * (unsigned)(ir & (XW - 1)), G.xpair + 2 * ir);
                }
                if (qr == nqw - 1) cpa_8(pt_lo + 256u * (unsigned)sdr, gw - planeB + 32);             /* OLD first point above the slab */
            }
            cp_async_commit();
        }

        /* (q1, i1) = position pos + 1 */
        const bool wrap1 = (q + 1 == nq);
        const int q1 = wrap1 ? 0 : q + 1, i1 = wrap1 ? i + 1 : i;
        char *gs1 = gs + 32 + (wrap1 ? wrapB : 0LL);
        const bool live = col_ok && (unsigned)i < (unsigned)nx && q < nqw;

        /* the groups of the last DP steps may still be in flight; everything this step reads was requested >= DP steps ago.
         * The barrier makes the landed copies and last step's NEW quads visible across the lanes of the warp. */
        if (DP == 2) cp_async_wait<2>(); else if (DP == 3) cp_async_wait<3>(); else cp_async_wait<1>();
        __syncwarp();

        /* ---- (D) this step's quad ---- */
        unsigned a1 = aw + GSS_SLOT; a1 = (a1 >= win_hi) ? a1 - RWB : a1;
        if (live) {
            /* ---- (C) entering a plane: x coefficients, reciprocal diagonal, -z value; at the top quad: the OLD point above ---- */
            if (q == 0) {
                const unsigned xa = xtab + 16u * (unsigned)(i & (XW - 1));
                cx2 = L.hyhz * lds_f64(xa); cx1 = L.hyhz * lds_f64(xa + 8u);
                s_xy = cx2 + cy2;
                if (ZUNI) { den_c = 4.0 * (s_xy + cz2_0); rden_c = lds_f64(rd_lo + 256u * (unsigned)sdi); }
                pB = has_below_cta ? lds_vf64(zb_mine + (unsigned)(i & (GSS_ZD - 1)) * 256u + (unsigned)b * 8u) : lds_f64(pb_lo + 512u * (unsigned)sdi + 8u);
            }
            if (q == nqw - 1) pTop = lds_f64(pt_lo + 256u * (unsigned)sdi);
            unsigned ap = aw + offP; ap = (ap >= win_hi) ? ap - RWB : ap;
            unsigned am = aw + (RWB - offP); am = (am >= win_hi) ? am - RWB : am;
            /* the x / y part of the numerator in two phases (the shared-memory loads are ordered asm statements: at most
             * two neighbour quads are live at a time, which keeps the kernel free of register spills -- with the whole
             * shared-memory carve-out in use the L1 is too small to hold spilled registers) */
            double txy[4];
            {
                const quad wp = lds_q(ap, GSS_HALF);                        /* OLD, plane i+1 */
                const quad wm = lds_q(am, GSS_HALF);                        /* NEW, plane i-1 (the x pad plane for i = 0) */
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    double u = cx2 * (wp.v[e] + wm.v[e]) * 2.0;
                    txy[e] = u + cx1 * (wp.v[e] - wm.v[e]) * L.hx;
                }
            }
            {
                const quad ym = lds_q(aw - 16u, GSS_HALF);                  /* NEW: lane b-1 wrote it one step ago (lane 0: halo cell) */
                const quad yp = lds_q(aw + 16u, GSS_HALF);                  /* OLD: lane b+1 updates it in the next step (last lane: halo cell) */
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    double u = txy[e] + cy2 * (yp.v[e] + ym.v[e]) * 2.0;
                    txy[e] = u + cy1 * (yp.v[e] - ym.v[e]) * L.hy;
                }
            }
            const quad rq = lds_q(ar, 512u);
            quad tz2, tz1;                              /* cz2, cz1 of the quad's four points */
            if (ZUNI) {                                 /* one value each along z: no table access */
#pragma unroll
                for (int e = 0; e < 4; e++) { tz2.v[e] = cz2_0; tz1.v[e] = cz1_0; }
            } else {
                tz2 = lds_q(ztab + 32u * q, 16u);
                tz1 = lds_q(ztab + 32u * (nq + q), 16u);
            }
            const quad cur = lds_q(aw, GSS_HALF);                           /* the quad's own OLD values */
            double pT3 = pTop;                                              /* +z of the quad's last point: above the slab, or ... */
            if (q != nqw - 1) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(pT3) : "r"(a1) : "memory");    /* ... the first point of position pos + 1 */
            quad o;
            if (q != qpart) {
                gss_chain<ZUNI, false>(txy, cur, pT3, rq, tz2, tz1, L.hz, G.two_h, s_xy, den_c, rden_c, 4, pB, o);
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(gs), "d"(o.v[0]), "d"(o.v[1]), "d"(o.v[2]), "d"(o.v[3]) : "memory");
            } else {
                gss_chain<ZUNI, true>(txy, cur, pT3, rq, tz2, tz1, L.hz, G.two_h, s_xy, den_c, rden_c, nvalid_last, pB, o);
                double *dst = (double *)gs;
                dst[0] = o.v[0];
                if (nvalid_last > 1) dst[1] = o.v[1];
                if (nvalid_last > 2) dst[2] = o.v[2];
            }
            sts_q(aw, GSS_HALF, o);                      /* in place: the NEW quad stays in the window for lane b+1 and for plane i+1 */
            if (has_above_cta && q == nq - 1)
                asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(zb_above + (unsigned)(i & (GSS_ZD - 1)) * 256u + (unsigned)b * 8u), "d"(o.v[3]) : "memory");
        }

        /* ---- (E) advance ---- */
        q = q1; i = i1; gs = gs1;
        if (wrap1) sdi = (sdi + 1 == SD) ? 0 : sdi + 1;
        const bool wrapr = (++qr == nq);
        gw += 32 + (wrapr ? wrapB : 0LL);
        if (wrapr) { qr = 0; ir += 1; sdr = (sdr + 1 == SD) ? 0 : sdr + 1; }
        aw = a1;
        ar += GSS_RSLOT; ar = (ar >= rr_hi) ? ar - RRB : ar;
        __syncwarp();
#ifdef GSS_EXPERIMENT_NO_FENCE      /* timing experiment only: NOT safe */
        if (b == 0 && t >= 0) s_done[a] = t + 1;
#else
        if (b == 0 && t >= 0) { gss_fence_cta(); s_done[a] = t + 1; }
#endif
    }
}
