/*
 * mg_tma.cuh -- TMA-staged x-marching stencil kernel for the fine 3-D levels (Jacobi sweep d3:153-166, residual
 * d3:96-108, error norm d3:408-430).  Same arithmetic, operand for operand, as k_stencil_march / k_error_march in
 * mg_kernels.cuh (the point formulas are shared); what changes is how the operands reach the SM.
 *
 * The register-marching kernels keep one x-plane of loads in flight per thread, so the bytes in flight per SM are bounded
 * by occupancy (~32 KB) and the sweep is latency-bound at ~73 % of the measured copy bandwidth.  Here a CTA owns a
 * (TY x TZ) tile of (y, z) and marches x.  One elected thread streams the tile of every x-plane -- pressure with its
 * one-point halo ring, and the right-hand side -- into a ring of shared-memory stages with the Tensor Memory Accelerator
 * (cp.async.bulk.tensor.3d, completion on an mbarrier, out-of-range coordinates zero-filled by the hardware), MG_TMA_STAGES-1
 * planes ahead of the arithmetic.  Bytes in flight are now set by the ring (2 CTAs x 4 planes x 18.4 KB per SM), not by
 * registers.  A thread owns the aligned z pair (k0, k0+1) of two adjacent rows (four points per x step: the per-step
 * overhead -- barrier wait, addressing, loop -- is paid once for four points, and the rows exchange their y neighbours in
 * registers); the x-direction neighbours stay in registers as before, the remaining y / z neighbours and the next plane
 * come from shared memory with conflict-free 128-bit loads; results go straight to global memory with 128-bit stores.
 * With the operands staged this way the kernel is bound by instruction issue, not by memory latency, which is why the
 * instruction count per point matters (ncu: profiles/r01_ncu_ops_513_tma*.txt).
 *
 * Tile geometry in the level array (z fastest, off(i,j,k) = (i+1)*plane + (j+1)*pitch + k + 4):
 *   pressure box  = [TZ+4, TY+2, 1] starting at (k0-2, j0-1, i): column c of a tile row holds k = k0 - 2 + c, so the
 *                   thread's own pair sits on a 16-byte boundary (c = 2*tx + 2);
 *   rhs box       = [TZ, TY, 1] starting at (k0, j0, i).
 */
#pragma once
#include <cuda.h>
#include <cstdint>

#include "mg_kernels.cuh"

#ifndef MG_TMA_STAGES
#define MG_TMA_STAGES 5
#endif
#define MG_TMA_TZ 128
#define MG_TMA_TY 8
#define MG_TMA_WARPS (MG_TMA_TZ / 2 * (MG_TMA_TY / 2) / 32)         /* 8: a thread owns a z pair of TWO adjacent rows */
#define MG_TMA_PW (MG_TMA_TZ + 4)                                   /* doubles per pressure tile row */
#define MG_TMA_P_BYTES ((MG_TMA_TY + 2) * MG_TMA_PW * 8)            /* 10560 */
#define MG_TMA_P_SLOT (((MG_TMA_P_BYTES + 127) / 128) * 128)        /* 10624 */
#define MG_TMA_R_BYTES (MG_TMA_TY * MG_TMA_TZ * 8)                  /* 8192 */
#define MG_TMA_STAGE_BYTES (MG_TMA_P_SLOT + MG_TMA_R_BYTES)         /* 18816 */
#define MG_TMA_STAGE_D (MG_TMA_STAGE_BYTES / 8)                     /* doubles per stage */
#define MG_TMA_MAXCH 384                                            /* x planes per CTA at most (x-coefficient table) */
#define MG_TMA_SMEM (MG_TMA_STAGES * MG_TMA_STAGE_BYTES)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(const uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const uint32_t dst, const CUtensorMap *map, const uint32_t bar, const int c0, const int c1, const int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

/* MODE 0 Jacobi (FAST = diagonal invariant along x), 1 residual (FAST = power-of-two divisors), 2 error norm (FAST = pow2).
 * part: MODE 2 only, [0*nb + bid] = num, [1*nb + bid] = sum rhs^2 (layout of k_error_march).
 *
 * Pipeline: plane sequence n = 0 .. nplanes-1 is x-plane i0-1+n and lives in ring stage n % STAGES.  full[s] (count 1 +
 * transaction bytes) completes when the TMA has filled stage s; empty[s] (count = warps) completes when every warp has
 * finished reading it.  Thread 0 refills the stage released one iteration earlier, so warps drift up to a plane apart
 * without ever meeting at a CTA-wide barrier. */
struct TmaRow {                 /* per-row state of a thread */
    double2 pW, pC;             /* own pair on planes i-1, i */
    double cy2, cy1, y2, y1;    /* MODE 0: hzhx*ety2, hzhx*etyy; MODE 1/2: ety2, etyy */
    double den_a, den_b, rden_a, rden_b;
};

template <int MODE, bool FAST>
__device__ __forceinline__ void tma_point_pair(const TmaRow &r, const double2 pE, const double2 pS, const double2 pN, const double pBa,
                                               const double pTb, const double2 rr, const double xa, const double xb,
                                               const double hx, const double hy, const double hz, const double two_h,
                                               const double cz2a, const double cz2b, const double cz1a, const double cz1b,
                                               const double z2a, const double z2b, const double z1a, const double z1b, const LapDiv &dv,
                                               const bool has2, const int norm_order, double2 &out, double &num, double &den) {
    const double2 pC = r.pC, pW = r.pW;
    if (MODE == 0) {
        if (FAST) {
            const double ta = relax3_num(pE.x, pW.x, pN.x, pS.x, pC.y, pBa, rr.x, xa, xb, hx, r.cy2, r.cy1, hy, cz2a, cz1a, hz, two_h);
            const double tb = relax3_num(pE.y, pW.y, pN.y, pS.y, pTb, pC.x, rr.y, xa, xb, hx, r.cy2, r.cy1, hy, cz2b, cz1b, hz, two_h);
            out.x = div_inv_fast(ta, r.den_a, r.rden_a);
            out.y = div_inv_fast(tb, r.den_b, r.rden_b);
            if (div_inv_bad(ta) | div_inv_bad(tb)) {                /* rare: outside the range where the fast sequence is exact */
                out.x = div_ieee(ta, r.den_a);
                out.y = div_ieee(tb, r.den_b);
            }
        } else {
            const double s = xa + r.cy2;
            out.x = relax3(pE.x, pW.x, pN.x, pS.x, pC.y, pBa, rr.x, xa, xb, hx, r.cy2, r.cy1, hy, cz2a, cz1a, hz, two_h, 4.0 * (s + cz2a));
            out.y = relax3(pE.y, pW.y, pN.y, pS.y, pTb, pC.x, rr.y, xa, xb, hx, r.cy2, r.cy1, hy, cz2b, cz1b, hz, two_h, 4.0 * (s + cz2b));
        }
    } else if (MODE == 1) {
        out.x = rr.x - lap3_inv<FAST>(pC.x, pE.x, pW.x, pN.x, pS.x, pC.y, pBa, xa, xb, r.y2, r.y1, z2a, z1a, dv);
        out.y = rr.y - lap3_inv<FAST>(pC.y, pE.y, pW.y, pN.y, pS.y, pTb, pC.x, xa, xb, r.y2, r.y1, z2b, z1b, dv);
    } else {
        const double ta = fabs(lap3_inv<FAST>(pC.x, pE.x, pW.x, pN.x, pS.x, pC.y, pBa, xa, xb, r.y2, r.y1, z2a, z1a, dv) - rr.x);
        const double tb = fabs(lap3_inv<FAST>(pC.y, pE.y, pW.y, pN.y, pS.y, pTb, pC.x, xa, xb, r.y2, r.y1, z2b, z1b, dv) - rr.y);
        if (norm_order == 1) {
            num = fmax(num, ta);
            if (has2) num = fmax(num, tb);
        } else {
            num += ta * ta; den += rr.x * rr.x;
            if (has2) { num += tb * tb; den += rr.y * rr.y; }
        }
    }
}

template <int MODE, bool FAST, bool PUSH>
__device__ __forceinline__ void stencil_tma_body(const CUtensorMap *mapP, const CUtensorMap *mapR, double *__restrict__ O,
                                                 const LevelView &L, const int ichunk, const int norm_order, const double hx_c, const double hy_c,
                                                 const double hz_c, double *__restrict__ part, const int nblocks, const HaloPush *H) {
    extern __shared__ __align__(128) double tma_ring[];
    __shared__ __align__(8) uint64_t bars[2 * MG_TMA_STAGES];       /* full[0..S), empty[S..2S) */
    __shared__ __align__(16) double2 xt[MG_TMA_MAXCH];              /* per x-plane coefficient pair of this chunk */
    __shared__ double red_sm[32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int k_t = blockIdx.x * MG_TMA_TZ, j_t = blockIdx.y * MG_TMA_TY;
    const int k0 = k_t + 2 * tx, ja = j_t + 2 * ty;                 /* rows ja and ja+1 */
    const int i0 = blockIdx.z * ichunk;
    const int i1 = min(i0 + ichunk, L.nx);
    const int nloop = i1 - i0;
    const int nplanes = nloop + 2;                      /* planes i0-1 .. i1 */
    const bool act_a = (k0 < L.nz) && (ja < L.ny), act_b = (k0 < L.nz) && (ja + 1 < L.ny);
    const bool has2 = (k0 + 1 < L.nz);
    const bool producer = (tx == 0 && ty == 0);
    const bool lane0 = ((tx + ty * (MG_TMA_TZ / 2)) & 31) == 0;
    const uint32_t ring = smem_u32(tma_ring), bar0 = smem_u32(bars);

    if (producer) {
        for (int s = 0; s < MG_TMA_STAGES; s++) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (MG_TMA_STAGES + s), MG_TMA_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   /* x-plane table: MODE 0: kx*xix2[i], kx*xixx[i] (the leading products of the x terms); MODE 1/2: xix2[i], xixx[i] */
        const double kx = (MODE == 0) ? L.hyhz : 1.0;
        for (int q = tx + ty * (MG_TMA_TZ / 2); q < nloop; q += MG_TMA_WARPS * 32) {
            const double a = L.xix2[i0 + q], b = L.xixx[i0 + q];
            xt[q] = (MODE == 0) ? make_double2(kx * a, kx * b) : make_double2(a, b);
        }
    }
    __syncthreads();
    /* producer: rhs is needed for planes i0 .. i1-1 only */
    auto issue = [&](const int n, const int s) {
        const uint32_t st = ring + s * MG_TMA_STAGE_BYTES, bar = bar0 + 8 * s;
        const bool with_r = (n >= 1 && n <= nplanes - 2);
        mbar_expect_tx(bar, MG_TMA_P_BYTES + (with_r ? MG_TMA_R_BYTES : 0));
        tma_load_3d(st, mapP, bar, k_t + 2, j_t, i0 + n);                       /* (k_t-2)+4, (j_t-1)+1, (i0-1+n)+1 */
        if (with_r) tma_load_3d(st + MG_TMA_P_SLOT, mapR, bar, k_t + 4, j_t + 1, i0 + n);
    };
    if (producer)
        for (int n = 0; n < MG_TMA_STAGES && n < nplanes; n++) issue(n, n);

    /* per-thread constants (k0+1 may be the k = nz pad: its metric entries are 0 and its results are never stored) */
    double z2a = 0.0, z2b = 0.0, z1a = 0.0, z1b = 0.0;
    if (act_a) { z2a = L.ztz2[k0]; z2b = L.ztz2[k0 + 1]; z1a = L.ztzz[k0]; z1b = L.ztzz[k0 + 1]; }
    const double hx = L.hx, hy = L.hy, hz = L.hz;
    const double cz2a = L.hxhy * z2a, cz2b = L.hxhy * z2b, cz1a = L.hxhy * z1a, cz1b = L.hxhy * z1b;
    const double two_h = 2.0 * L.hxhyhz;
    LapDiv dv;
    if (MODE == 1) dv = make_lapdiv(hx * hx, 2.0 * hx, hy * hy, 2.0 * hy, hz * hz, 2.0 * hz, true);
    if (MODE == 2) dv = make_lapdiv(hx_c * hx_c, 2.0 * hx_c, hy_c * hy_c, 2.0 * hy_c, hz_c * hz_c, 2.0 * hz_c, true);
    TmaRow ra, rb;
    {
        const double e2a = act_a ? L.ety2[ja] : 0.0, e1a = act_a ? L.etyy[ja] : 0.0;
        const double e2b = act_b ? L.ety2[ja + 1] : 0.0, e1b = act_b ? L.etyy[ja + 1] : 0.0;
        ra.y2 = e2a; ra.y1 = e1a; rb.y2 = e2b; rb.y1 = e1b;
        ra.cy2 = L.hzhx * e2a; ra.cy1 = L.hzhx * e1a; rb.cy2 = L.hzhx * e2b; rb.cy1 = L.hzhx * e1b;
        ra.den_a = ra.den_b = rb.den_a = rb.den_b = 1.0;
        ra.rden_a = ra.rden_b = rb.rden_a = rb.rden_b = 1.0;
        if (MODE == 0 && FAST) {
            const double sa = xt[0].x + ra.cy2, sb = xt[0].x + rb.cy2;
            ra.den_a = 4.0 * (sa + cz2a); ra.den_b = 4.0 * (sa + cz2b);
            rb.den_a = 4.0 * (sb + cz2a); rb.den_b = 4.0 * (sb + cz2b);
            ra.rden_a = 1.0 / ra.den_a; ra.rden_b = 1.0 / ra.den_b; rb.rden_a = 1.0 / rb.den_a; rb.rden_b = 1.0 / rb.den_b;
        }
    }
    const int pa_off = (2 * ty + 1) * MG_TMA_PW + 2 * tx + 2;   /* own pair of row ja in a pressure tile (doubles); row ja+1 is PW further */
    const int rr_off = MG_TMA_P_SLOT / 8 + 2 * ty * MG_TMA_TZ + 2 * tx;
    double *ob = O + (long long)(ja + 1) * L.pitch + (k0 + 1 + MG_LEAD) + (long long)(i0 + 1) * L.plane;
    const long long plane = L.plane;
    const int pitch = L.pitch;
    double num = 0.0, den = 0.0;
    /* halo push (PUSH): loop iterations n whose x plane is a neighbour's pad, and whether one of this thread's rows is */
    int push_n0 = -1, push_n1 = -1;
    bool push_rows = false;
    if (PUSH) {
        if (H->dst[0]) push_n0 = 2 - i0;                    /* plane i = 1 is iteration n = i - i0 + 1 */
        if (H->dst[1]) push_n1 = L.nx - 1 - i0;             /* plane i = nx - 2 */
        push_rows = (H->dst[2] && (ja == 1 || ja + 1 == 1)) || (H->dst[3] && (ja == L.ny - 2 || ja + 1 == L.ny - 2));
    }

    /* prologue: own pairs of planes i0-1 (n = 0) and i0 (n = 1); stage 0 is released right away */
    mbar_wait(bar0, 0);
    ra.pW = *reinterpret_cast<const double2 *>(tma_ring + pa_off);
    rb.pW = *reinterpret_cast<const double2 *>(tma_ring + pa_off + MG_TMA_PW);
    mbar_wait(bar0 + 8, 0);
    ra.pC = *reinterpret_cast<const double2 *>(tma_ring + MG_TMA_STAGE_D + pa_off);
    rb.pC = *reinterpret_cast<const double2 *>(tma_ring + MG_TMA_STAGE_D + pa_off + MG_TMA_PW);
    __syncwarp();
    if (lane0) mbar_arrive(bar0 + 8 * MG_TMA_STAGES);

    int sc = 1, sn = 2 % MG_TMA_STAGES;                 /* stages of planes n and n+1 */
    int sp = 0;                                         /* stage of plane n-1: the one the producer refills in iteration n */
    uint32_t full_par = 0x3u & ((1u << MG_TMA_STAGES) - 1u);   /* bit s: parity of the NEXT wait on full[s] (stages 0, 1 consumed once) */
    uint32_t empty_par = 0;                             /* bit s: parity of the producer's next wait on empty[s] */
    for (int n = 1; n <= nloop; ++n) {
        mbar_wait(bar0 + 8 * sn, (full_par >> sn) & 1u);            /* plane n+1 has landed (plane n did earlier) */
        full_par ^= (1u << sn);
        const double *tc = tma_ring + sc * MG_TMA_STAGE_D + pa_off;
        const double *tn = tma_ring + sn * MG_TMA_STAGE_D + pa_off;
        const double2 pEa = *reinterpret_cast<const double2 *>(tn);
        const double2 pEb = *reinterpret_cast<const double2 *>(tn + MG_TMA_PW);
        if (act_a) {
            const double2 xc = xt[n - 1];
            const double2 pSa = *reinterpret_cast<const double2 *>(tc - MG_TMA_PW);
            const double2 pNb = *reinterpret_cast<const double2 *>(tc + 2 * MG_TMA_PW);
            const double pBa_a = reinterpret_cast<const double2 *>(tc - 2)->y, pTb_a = reinterpret_cast<const double2 *>(tc + 2)->x;
            const double pBa_b = reinterpret_cast<const double2 *>(tc + MG_TMA_PW - 2)->y, pTb_b = reinterpret_cast<const double2 *>(tc + MG_TMA_PW + 2)->x;
            const double *rt = tma_ring + sc * MG_TMA_STAGE_D + rr_off;
            const double2 rra = *reinterpret_cast<const double2 *>(rt);
            const double2 rrb = *reinterpret_cast<const double2 *>(rt + MG_TMA_TZ);
            double2 oa, obv;
            tma_point_pair<MODE, FAST>(ra, pEa, pSa, rb.pC, pBa_a, pTb_a, rra, xc.x, xc.y, hx, hy, hz, two_h, cz2a, cz2b, cz1a, cz1b,
                                       z2a, z2b, z1a, z1b, dv, has2, norm_order, oa, num, den);
            if (MODE != 2) { if (has2) st2(ob, oa); else ob[0] = oa.x; }
            if (act_b) {
                tma_point_pair<MODE, FAST>(rb, pEb, ra.pC, pNb, pBa_b, pTb_b, rrb, xc.x, xc.y, hx, hy, hz, two_h, cz2a, cz2b, cz1a, cz1b,
                                           z2a, z2b, z1a, z1b, dv, has2, norm_order, obv, num, den);
                if (MODE != 2) { if (has2) st2(ob + pitch, obv); else ob[pitch] = obv.x; }
            }
            if (PUSH && (push_rows | (n == push_n0) | (n == push_n1))) {         /* rare; kept out of the way: the loop is issue-bound */
                halo_push_store(*H, L, i0 + n - 1, ja, k0, oa, has2);
                if (act_b) halo_push_store(*H, L, i0 + n - 1, ja + 1, k0, obv, has2);
            }
        }
        ra.pW = ra.pC; ra.pC = pEa;
        rb.pW = rb.pC; rb.pC = pEb;
        ob += plane;
        __syncwarp();
        if (lane0) mbar_arrive(bar0 + 8 * (MG_TMA_STAGES + sc));    /* this warp is done with plane n */
        if (producer && n - 1 + MG_TMA_STAGES < nplanes) {
            mbar_wait(bar0 + 8 * (MG_TMA_STAGES + sp), (empty_par >> sp) & 1u);    /* every warp is done with plane n-1 */
            empty_par ^= (1u << sp);
            issue(n - 1 + MG_TMA_STAGES, sp);
        }
        sp = sc; sc = sn; sn = (sn + 1 == MG_TMA_STAGES) ? 0 : sn + 1;
    }
    if (MODE == 2) {
        const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        const double n2 = (norm_order == 1) ? block_reduce<true>(num, red_sm) : block_reduce<false>(num, red_sm);
        const double d2 = block_reduce<false>(den, red_sm);
        if (tx == 0 && ty == 0) { part[bid] = n2; part[nblocks + bid] = d2; }
    }
}

template <int MODE, bool FAST>
__global__ void __launch_bounds__(MG_TMA_WARPS * 32, 2)
k_stencil_tma(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapR, double *__restrict__ O,
              const LevelView L, const int ichunk, const int norm_order, const double hx_c, const double hy_c, const double hz_c,
              double *__restrict__ part, const int nblocks) {
    stencil_tma_body<MODE, FAST, false>(&mapP, &mapR, O, L, ichunk, norm_order, hx_c, hy_c, hz_c, part, nblocks, nullptr);
}
/* Jacobi sweep of a rank with neighbours: the same arithmetic plus the halo push (mg_kernels.cuh, HaloPush) */
template <bool FAST>
__global__ void __launch_bounds__(MG_TMA_WARPS * 32, 2)
k_stencil_tma_push(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapR, double *__restrict__ O,
                   const LevelView L, const int ichunk, const HaloPush H) {
    stencil_tma_body<0, FAST, true>(&mapP, &mapR, O, L, ichunk, 0, 0.0, 0.0, 0.0, nullptr, 0, &H);
    const int i0 = blockIdx.z * ichunk, j0 = blockIdx.y * MG_TMA_TY;
    halo_push_signal(H, halo_cta_pushes(H, L, i0, min(i0 + ichunk, L.nx), j0, j0 + MG_TMA_TY));
}
