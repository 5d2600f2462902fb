/*
 * mg_fused.cuh -- computeResidual (d3:90-111) fused with the 27-point full-weighting restriction (coarsen, d3:250-298) on the
 * fine 3-D levels: the residual never goes to HBM.
 *
 * Unfused, the pair costs 24 + 9 B per fine point (write tmpDataArray, read it back); fused it costs 16 + 1 (read pressure
 * and rhs, write the coarse rhs).  The arithmetic is unchanged operand for operand: the residual point formula is the one
 * of k_stencil_tma<1> (tma_point_pair), the restriction sums are those of k_coarsen_march in the order of d3:270-297.
 *
 * Split of the work.  A coarse point (I, J, K) reads the fine residual on the 3 x 3 x 3 box around (2I, 2J, 2K).
 *   INTERIOR coarse points (1 <= I <= ncx-2, ...) only touch fine CORE points 1 .. n-2, whose residual depends on nothing
 *     but core pressure values: no pad, no boundary condition, no halo.  They are k_resid_coarsen_tma's.
 *   BOUNDARY coarse points (a face of the coarse box) read pads of tmpDataArray, i.e. zeros at walls, wrapped values on
 *     periodic axes, neighbour values between ranks (updatePads, incl. the edge pads of SURVEY App. A-6).  For those the
 *     unfused sequence runs on a two-plane shell: k_resid_shell writes the residual of the fine points within two planes of
 *     a face into tmpDataArray, the usual zero-pads / updatePads follow, k_coarsen_shell restricts the coarse faces.
 *
 * k_resid_coarsen_tma.  A CTA owns CJ x 63 coarse (J, K) columns = an FY x 128 fine (y, z) footprint (2*CJ+1 x 127 needed;
 * neighbouring CTAs overlap by two fine rows / columns) and marches x over a chunk of coarse planes.  FY = 8 (CJ = 3, two
 * CTAs per SM, a third of the rows recomputed; the overlapping rows are L2 hits) or 16 (CJ = 7, one CTA per SM, 16 %).
 * Measured on B200 at 1025^3 (profiles/r02_fused_residual_restriction.txt): residual + restriction 6.06 ms unfused, 5.60 ms
 * with 16-row tiles, 7.55 ms with 8-row tiles; DRAM traffic 18.9 GB (1.03 x the 17 B/pt), issue slots 44 % and fp64 pipe 45 %
 * busy with 16 warps per SM -- the kernel is bound by per-warp dependent-issue latency, like k_stencil_tma<1> itself, so
 * the saved HBM traffic buys only 8 % and on smaller levels (tile quantisation in z) nothing: fused on levels of >= 4e8
 * points only (sarasmg.cu: fuse_level_ok).  Operands
 * are staged exactly as in k_stencil_tma: one thread streams the pressure tile (with halo ring) and the rhs tile of every
 * fine x-plane through a shared-memory ring with TMA, full / empty mbarriers, a thread owns an aligned z pair of two
 * adjacent rows.  Residual planes go into a second shared-memory ring of five planes; after every second plane the CTA
 * meets at one barrier and its first CJ x 63 threads each restrict one coarse point from the last three planes.  With five
 * slots a warp that runs ahead writes planes m+1, m+2 while a slow one still reads m-2 .. m: one barrier per coarse plane.
 */
#pragma once
#include "mg_tma.cuh"

#define MG_FZ 128                                                   /* fine columns per CTA */
#ifndef MG_FY
#define MG_FY 16                                                    /* fine rows per CTA (16: one CTA per SM; 8: two) */
#endif
#define MG_F_CJ ((MG_FY - 2) / 2)                                   /* coarse rows per CTA: 2*CJ + 1 <= FY */
#define MG_F_CK 63                                                  /* coarse columns per CTA: 2*CK + 1 <= FZ - 1 (the footprint starts one column early: alignment) */
#define MG_F_WARPS (MG_FZ / 2 * (MG_FY / 2) / 32)
#ifndef MG_F_STAGES
#define MG_F_STAGES (MG_FY == 8 ? 3 : 4)
#endif
#define MG_F_MINB (MG_FY == 8 ? 2 : 1)
#define MG_F_PW (MG_FZ + 4)
#define MG_F_P_BYTES ((MG_FY + 2) * MG_F_PW * 8)
#define MG_F_P_SLOT (((MG_F_P_BYTES + 127) / 128) * 128)
#define MG_F_R_BYTES (MG_FY * MG_FZ * 8)
#define MG_F_STAGE_BYTES (MG_F_P_SLOT + MG_F_R_BYTES)
#define MG_F_STAGE_D (MG_F_STAGE_BYTES / 8)
#define MG_F_RING 5                                                 /* residual planes kept */
#define MG_F_PLANE_D (MG_FY * MG_FZ)                                /* doubles per residual plane */
#define MG_F_MAXCI 63                                               /* coarse planes per CTA at most: 2*CI + 1 fine planes in the x table */
#define MG_F_SMEM (MG_F_STAGES * MG_F_STAGE_BYTES + MG_F_RING * MG_F_PLANE_D * 8)       /* 8 rows: 97408; 16 rows: 223744 */

/* FAST: the six divisors of the residual are powers of two.  CI: coarse planes per x chunk. */
template <bool FAST>
__global__ void __launch_bounds__(MG_F_WARPS * 32, MG_F_MINB)
k_resid_coarsen_tma(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapR, const LevelView L,
                    double *__restrict__ Rc, const LevelView Lc, const int CI) {
    extern __shared__ __align__(128) double fz_smem[];
    __shared__ __align__(8) uint64_t bars[2 * MG_F_STAGES];
    __shared__ __align__(16) double2 xt[2 * MG_F_MAXCI + 2];
    double *ring_d = fz_smem;                                           /* operand stages */
    double *res = fz_smem + MG_F_STAGES * MG_F_STAGE_D;                 /* residual planes */
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = tx + ty * (MG_FZ / 2);
    const int k_t = blockIdx.x * (2 * MG_F_CK);                         /* first fine column of the footprint (even) */
    const int j_t = blockIdx.y * (2 * MG_F_CJ) + 1;                     /* first fine row */
    const int k0 = k_t + 2 * tx, ja = j_t + 2 * ty;
    const int Ia = 1 + blockIdx.z * CI;                                 /* coarse planes Ia .. Ib */
    const int Ib = min(Ia + CI - 1, Lc.nx - 2);
    const int i0 = 2 * Ia - 1;                                          /* first fine plane whose residual is needed */
    const int nloop = 2 * (Ib - Ia) + 3;                                /* fine residual planes i0 .. 2*Ib + 1 */
    const int nplanes = nloop + 2;
    const bool act_a = (k0 < L.nz) && (ja < L.ny), act_b = (k0 < L.nz) && (ja + 1 < L.ny);
    const bool producer = (tid == 0);
    const bool lane0 = (tid & 31) == 0;
    const uint32_t ring = smem_u32(ring_d), bar0 = smem_u32(bars);

    if (producer) {
        for (int s = 0; s < MG_F_STAGES; s++) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (MG_F_STAGES + s), MG_F_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int q = tid; q < nloop; q += MG_F_WARPS * 32) xt[q] = make_double2(L.xix2[i0 + q], L.xixx[i0 + q]);
    __syncthreads();
    auto issue = [&](const int n, const int s) {
        const uint32_t st = ring + s * MG_F_STAGE_BYTES, bar = bar0 + 8 * s;
        const bool with_r = (n >= 1 && n <= nplanes - 2);
        mbar_expect_tx(bar, MG_F_P_BYTES + (with_r ? MG_F_R_BYTES : 0));
        tma_load_3d(st, &mapP, bar, k_t + 2, j_t, i0 + n);                      /* (k_t-2)+4, (j_t-1)+1, (i0-1+n)+1 */
        if (with_r) tma_load_3d(st + MG_F_P_SLOT, &mapR, bar, k_t + 4, j_t + 1, i0 + n);
    };
    if (producer)
        for (int n = 0; n < MG_F_STAGES && n < nplanes; n++) issue(n, n);

    double z2a = 0.0, z2b = 0.0, z1a = 0.0, z1b = 0.0;
    if (act_a) { z2a = L.ztz2[k0]; z2b = L.ztz2[k0 + 1]; z1a = L.ztzz[k0]; z1b = L.ztzz[k0 + 1]; }
    const double hx = L.hx, hy = L.hy, hz = L.hz;
    const LapDiv dv = make_lapdiv(hx * hx, 2.0 * hx, hy * hy, 2.0 * hy, hz * hz, 2.0 * hz, true);
    TmaRow ra, rb;
    ra.y2 = act_a ? L.ety2[ja] : 0.0; ra.y1 = act_a ? L.etyy[ja] : 0.0;
    rb.y2 = act_b ? L.ety2[ja + 1] : 0.0; rb.y1 = act_b ? L.etyy[ja + 1] : 0.0;
    ra.cy2 = ra.cy1 = rb.cy2 = rb.cy1 = 0.0;
    ra.den_a = ra.den_b = rb.den_a = rb.den_b = ra.rden_a = ra.rden_b = rb.rden_a = rb.rden_b = 1.0;
    const int pa_off = (2 * ty + 1) * MG_F_PW + 2 * tx + 2;
    const int rr_off = MG_F_P_SLOT / 8 + 2 * ty * MG_FZ + 2 * tx;
    const int my_res = 2 * ty * MG_FZ + 2 * tx;                         /* own pair of row ja in a residual plane */
    double num = 0.0, den = 0.0;                                        /* unused outputs of tma_point_pair<1> */

    /* restriction role of this thread: coarse column (J, K) of the tile, rows 2*cj .. 2*cj+2, columns 2*ck+1 .. 2*ck+3 */
    const int cj = tid / MG_F_CK, ck = tid - cj * MG_F_CK;
    const int J = 1 + blockIdx.y * MG_F_CJ + cj, K = 1 + blockIdx.x * MG_F_CK + ck;
    const bool coarse_act = (cj < MG_F_CJ) && (J <= Lc.ny - 2) && (K <= Lc.nz - 2);
    const int c_off = (2 * cj) * MG_FZ + 2 * ck + 1;                    /* element (row 2cj, column 2ck+1) */
    double *rc = Rc + (long long)(J + 1) * Lc.pitch + (K + 1 + MG_LEAD);

    mbar_wait(bar0, 0);
    ra.pW = *reinterpret_cast<const double2 *>(ring_d + pa_off);
    rb.pW = *reinterpret_cast<const double2 *>(ring_d + pa_off + MG_F_PW);
    mbar_wait(bar0 + 8, 0);
    ra.pC = *reinterpret_cast<const double2 *>(ring_d + MG_F_STAGE_D + pa_off);
    rb.pC = *reinterpret_cast<const double2 *>(ring_d + MG_F_STAGE_D + pa_off + MG_F_PW);
    __syncwarp();
    if (lane0) mbar_arrive(bar0 + 8 * MG_F_STAGES);

    int sc = 1, sn = 2 % MG_F_STAGES, sp = 0;
    uint32_t full_par = 0x3u & ((1u << MG_F_STAGES) - 1u);
    uint32_t empty_par = 0;
    int slot = 0;                                                       /* residual-ring slot of plane m = n - 1 */
    for (int n = 1; n <= nloop; ++n) {
        mbar_wait(bar0 + 8 * sn, (full_par >> sn) & 1u);
        full_par ^= (1u << sn);
        const double *tc = ring_d + sc * MG_F_STAGE_D + pa_off;
        const double *tn = ring_d + sn * MG_F_STAGE_D + pa_off;
        const double2 pEa = *reinterpret_cast<const double2 *>(tn);
        const double2 pEb = *reinterpret_cast<const double2 *>(tn + MG_F_PW);
        if (act_a) {
            const double2 xc = xt[n - 1];
            const double2 pSa = *reinterpret_cast<const double2 *>(tc - MG_F_PW);
            const double2 pNb = *reinterpret_cast<const double2 *>(tc + 2 * MG_F_PW);
            const double pBa_a = reinterpret_cast<const double2 *>(tc - 2)->y, pTb_a = reinterpret_cast<const double2 *>(tc + 2)->x;
            const double pBa_b = reinterpret_cast<const double2 *>(tc + MG_F_PW - 2)->y, pTb_b = reinterpret_cast<const double2 *>(tc + MG_F_PW + 2)->x;
            const double *rt = ring_d + sc * MG_F_STAGE_D + rr_off;
            const double2 rra = *reinterpret_cast<const double2 *>(rt);
            const double2 rrb = *reinterpret_cast<const double2 *>(rt + MG_FZ);
            double2 oa, obv;
            tma_point_pair<1, FAST>(ra, pEa, pSa, rb.pC, pBa_a, pTb_a, rra, xc.x, xc.y, hx, hy, hz, 0.0, 0.0, 0.0, 0.0, 0.0,
                                    z2a, z2b, z1a, z1b, dv, true, 0, oa, num, den);
            double *rp = res + slot * MG_F_PLANE_D + my_res;
            *reinterpret_cast<double2 *>(rp) = oa;
            if (act_b) {
                tma_point_pair<1, FAST>(rb, pEb, ra.pC, pNb, pBa_b, pTb_b, rrb, xc.x, xc.y, hx, hy, hz, 0.0, 0.0, 0.0, 0.0, 0.0,
                                        z2a, z2b, z1a, z1b, dv, true, 0, obv, num, den);
                *reinterpret_cast<double2 *>(rp + MG_FZ) = obv;
            }
        }
        ra.pW = ra.pC; ra.pC = pEa;
        rb.pW = rb.pC; rb.pC = pEb;
        __syncwarp();
        if (lane0) mbar_arrive(bar0 + 8 * (MG_F_STAGES + sc));
        if (producer && n - 1 + MG_F_STAGES < nplanes) {
            mbar_wait(bar0 + 8 * (MG_F_STAGES + sp), (empty_par >> sp) & 1u);
            empty_par ^= (1u << sp);
            issue(n - 1 + MG_F_STAGES, sp);
        }
        sp = sc; sc = sn; sn = (sn + 1 == MG_F_STAGES) ? 0 : sn + 1;

        if (n >= 3 && (n & 1)) {                        /* residual planes m = n-3, n-2, n-1 are fine planes 2I-1, 2I, 2I+1 */
            __syncthreads();
            if (coarse_act) {
                const int I = Ia + (n - 3) / 2;
                const int sC = slot, sB = (slot + MG_F_RING - 1) % MG_F_RING, sA = (slot + MG_F_RING - 2) % MG_F_RING;
                const double *a = res + sA * MG_F_PLANE_D + c_off, *b = res + sB * MG_F_PLANE_D + c_off, *c = res + sC * MG_F_PLANE_D + c_off;
                double A[3][3], B[3][3], C[3][3];
#pragma unroll
                for (int r = 0; r < 3; r++) {           /* column 2ck+1, then the aligned pair (2ck+2, 2ck+3) */
                    const double2 va = *reinterpret_cast<const double2 *>(a + r * MG_FZ + 1);
                    const double2 vb = *reinterpret_cast<const double2 *>(b + r * MG_FZ + 1);
                    const double2 vc = *reinterpret_cast<const double2 *>(c + r * MG_FZ + 1);
                    A[r][0] = a[r * MG_FZ]; A[r][1] = va.x; A[r][2] = va.y;
                    B[r][0] = b[r * MG_FZ]; B[r][1] = vb.x; B[r][2] = vb.y;
                    C[r][0] = c[r * MG_FZ]; C[r][1] = vc.x; C[r][2] = vc.y;
                }
                const double facePoints = (C[1][1] + A[1][1] + B[2][1] + B[0][1] + B[1][2] + B[1][0]) * 0.0625;
                const double edgePoints = (C[2][1] + C[0][1] + A[0][1] + A[2][1] + B[2][2] + B[0][2] + B[0][0] + B[2][0] +
                                           C[1][2] + C[1][0] + A[1][0] + A[1][2]) * 0.03125;
                const double vertPoints = (C[2][2] + C[2][0] + C[0][2] + A[2][2] + C[0][0] + A[2][0] + A[0][2] + A[0][0]) * 0.015625;
                rc[(long long)(I + 1) * Lc.plane] = facePoints + edgePoints + vertPoints + B[1][1] * 0.125;
            }
        }
        slot = (slot + 1 == MG_F_RING) ? 0 : slot + 1;
    }
}

/* ---- the two-plane shell ---- */
/* fine core points within two planes of a face of the rank's box, enumerated without repetition:
 *   x slabs  i in {0, 1, nx-2, nx-1}, all (j, k);   y slabs  j in {..}, 2 <= i <= nx-3, all k;   z slabs  k in {..}, i and j interior */
__device__ __forceinline__ int shell_edge_index(const int q, const int n) { return q < 2 ? q : n - 4 + q; }      /* q = 0..3 -> 0, 1, n-2, n-1 */
template <bool FAST>
__global__ void __launch_bounds__(256) k_resid_shell(const double *__restrict__ P, const double *__restrict__ R, double *__restrict__ T, const LevelView L) {
    const long long nX = 4LL * L.ny * L.nz;
    const long long nY = 4LL * (L.nx - 4) * L.nz;
    const long long nZ = 4LL * (L.nx - 4) * (L.ny - 4);
    const double hx = L.hx, hy = L.hy, hz = L.hz;
    const LapDiv dv = make_lapdiv(hx * hx, 2.0 * hx, hy * hy, 2.0 * hy, hz * hz, 2.0 * hz, true);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nX + nY + nZ; q += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (q < nX) {
            k = (int)(q % L.nz); const long long w = q / L.nz;
            j = (int)(w % L.ny); i = shell_edge_index((int)(w / L.ny), L.nx);
        } else if (q < nX + nY) {
            const long long u = q - nX;
            k = (int)(u % L.nz); const long long w = u / L.nz;
            i = 2 + (int)(w % (L.nx - 4)); j = shell_edge_index((int)(w / (L.nx - 4)), L.ny);
        } else {
            const long long u = q - nX - nY;
            k = shell_edge_index((int)(u & 3), L.nz); const long long w = u >> 2;
            j = 2 + (int)(w % (L.ny - 4)); i = 2 + (int)(w / (L.ny - 4));
        }
        const long long o = mg_at(L, i, j, k);
        const double *c = P + o;
        T[o] = R[o] - lap3_inv<FAST>(c[0], c[L.plane], c[-L.plane], c[L.pitch], c[-L.pitch], c[1], c[-1],
                                     L.xix2[i], L.xixx[i], L.ety2[j], L.etyy[j], L.ztz2[k], L.ztzz[k], dv);
    }
}

/* coarse points on a face of the coarse box (the others are k_resid_coarsen_tma's): restriction as in k_coarsen<3> */
__global__ void __launch_bounds__(256) k_coarsen_shell(const double *__restrict__ T, const LevelView Lf, double *__restrict__ Rc, const LevelView Lc) {
    const long long nX = 2LL * Lc.ny * Lc.nz;
    const long long nY = 2LL * (Lc.nx - 2) * Lc.nz;
    const long long nZ = 2LL * (Lc.nx - 2) * (Lc.ny - 2);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nX + nY + nZ; q += (long long)gridDim.x * blockDim.x) {
        int i, j, k;
        if (q < nX) {
            k = (int)(q % Lc.nz); const long long w = q / Lc.nz;
            j = (int)(w % Lc.ny); i = (w / Lc.ny) ? Lc.nx - 1 : 0;
        } else if (q < nX + nY) {
            const long long u = q - nX;
            k = (int)(u % Lc.nz); const long long w = u / Lc.nz;
            i = 1 + (int)(w % (Lc.nx - 2)); j = (w / (Lc.nx - 2)) ? Lc.ny - 1 : 0;
        } else {
            const long long u = q - nX - nY;
            k = (u & 1) ? Lc.nz - 1 : 0; const long long w = u >> 1;
            j = 1 + (int)(w % (Lc.ny - 2)); i = 1 + (int)(w / (Lc.ny - 2));
        }
        const int i2 = 2 * i, j2 = 2 * j, k2 = 2 * k;
#define TF(a, b, c) T[mg_at(Lf, (a), (b), (c))]
        const double facePoints = (TF(i2 + 1, j2, k2) + TF(i2 - 1, j2, k2) + TF(i2, j2 + 1, k2) + TF(i2, j2 - 1, k2) +
                                   TF(i2, j2, k2 + 1) + TF(i2, j2, k2 - 1)) * 0.0625;
        const double edgePoints = (TF(i2 + 1, j2 + 1, k2) + TF(i2 + 1, j2 - 1, k2) + TF(i2 - 1, j2 - 1, k2) + TF(i2 - 1, j2 + 1, k2) +
                                   TF(i2, j2 + 1, k2 + 1) + TF(i2, j2 - 1, k2 + 1) + TF(i2, j2 - 1, k2 - 1) + TF(i2, j2 + 1, k2 - 1) +
                                   TF(i2 + 1, j2, k2 + 1) + TF(i2 + 1, j2, k2 - 1) + TF(i2 - 1, j2, k2 - 1) + TF(i2 - 1, j2, k2 + 1)) * 0.03125;
        const double vertPoints = (TF(i2 + 1, j2 + 1, k2 + 1) + TF(i2 + 1, j2 + 1, k2 - 1) + TF(i2 + 1, j2 - 1, k2 + 1) + TF(i2 - 1, j2 + 1, k2 + 1) +
                                   TF(i2 + 1, j2 - 1, k2 - 1) + TF(i2 - 1, j2 + 1, k2 - 1) + TF(i2 - 1, j2 - 1, k2 + 1) + TF(i2 - 1, j2 - 1, k2 - 1)) * 0.015625;
        Rc[mg_at(Lc, i, j, k)] = facePoints + edgePoints + vertPoints + TF(i2, j2, k2) * 0.125;
#undef TF
    }
}
