/*
 * mg_kernels.cuh -- hand-written sm_100a fp64 kernels of the multigrid pressure-Poisson solver.
 *
 * Every kernel reproduces the arithmetic of one loop nest of SARAS's lib/poisson (cited per kernel,
 * "d3:N" = lib/poisson/poisson_d3.cc line N, "d2:N" = poisson_d2.cc, "p:N" = poisson.cc) operand for
 * operand: this translation unit is compiled with -fmad=false, every division stays an IEEE division,
 * and the association of every sum is the one the reference source spells out.  The path is an
 * HBM-bound 7-point stencil: no tensor cores; the work is coalescing (z is the unit-stride axis and
 * each thread owns an aligned pair of z points -> 128-bit loads/stores), register-blocked marching
 * along x (three x-planes of the thread's column live in registers, so every pressure value is
 * fetched from DRAM once per sweep) and keeping enough loads in flight.
 *
 * Device layout of a level array (private to the library): padded extents F = n + 2 with index base
 * -1 as in the reference, z rows padded to `pitch` doubles with MG_LEAD leading slack doubles so that
 * core point k = 0 of every row sits on a 32-byte sector boundary:
 *     off(i,j,k) = (i+1)*plane + (j+1)*pitch + (k + 1 + MG_LEAD),  plane = F_y*pitch,  pitch % 4 == 0.
 */
#pragma once
#include <cuda_runtime.h>

#define MG_LEAD 3

struct LevelView {
    int nx, ny, nz;         /* core points (xEnd+1, yEnd+1, zEnd+1) */
    int Fx, Fy, Fz;         /* padded extents */
    int pitch;              /* doubles per z row */
    long long plane;        /* doubles per x plane = Fy*pitch */
    /* metric arrays, pointers shifted so that [i] is index i (valid for -1..n) */
    const double *xixx, *xix2, *etyy, *ety2, *ztzz, *ztz2;
    double hx, hy, hz, hx2, hz2, hzhx, hxhy, hyhz, hxhyhz;
};

__host__ __device__ __forceinline__ long long mg_at(const LevelView &L, int i, int j, int k) {
    return (long long)(i + 1) * L.plane + (long long)(j + 1) * L.pitch + (long long)(k + 1 + MG_LEAD);
}

/* face-operation modes (imposeBC d3:575-668, updatePads self-neighbour d3:671-686, pad zeroing d3:91) */
enum { FACE_NONE = 0, FACE_MIRROR = 1, FACE_NEG_MIRROR = 2, FACE_WALL = 3, FACE_PERIODIC = 4, FACE_ZERO = 5 };

/* ------------------------------------------------------------------------------------------------ */
/* exact division by a loop-invariant divisor                                                         */
/* The reference divides by hx*hx, 2.0*hx, ... (6 divisions per residual point) and by the stencil     */
/* diagonal (1 per smoother point).  An IEEE fp64 division is ~12 dependent fp64-pipe instructions on  */
/* the GPU, which makes the residual compute-bound instead of HBM-bound.  When the divisor b does not  */
/* change along a thread's march, y = RN(1/b) is formed ONCE with a true IEEE division and every       */
/* quotient is then recovered exactly:  q0 = RN(a*y) is within 2 ulp of a/b, r = fma(-b, q0, a) is      */
/* exact, q1 = RN(q0 + r*y) is a faithful rounding of a/b, and one more step from a faithful quotient   */
/* with y = RN(1/b) returns RN(a/b) (Markstein 1990; Muller et al., Handbook of Floating-Point          */
/* Arithmetic, "division with an FMA").  The fused operations here do not contract any product of the   */
/* reference's expressions: they only rebuild the correctly rounded quotient the `/` would give, so     */
/* the result is bit-identical.  Exactness of r needs a*2^-106 to stay normal and a*y not to overflow:  */
/* operands outside 2^-823 <= |a| <= 2^777 take the IEEE division (a == 0 yields 0 on either path).     */
/* ------------------------------------------------------------------------------------------------ */
/* out of line on purpose: inlined, the compiler hoists the whole IEEE division sequence above the branch */
__device__ __noinline__ double div_ieee(const double a, const double b) { return a / b; }
__device__ __forceinline__ double div_inv(const double a, const double b, const double y) {
    double q = a * y;
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    const unsigned hi = (unsigned)__double2hiint(a);
    const unsigned e = (hi >> 20) & 0x7ffu;
    if ((e - 200u) > 1600u) {                                       /* rare: tiny, huge, zero, inf, nan */
        if (((hi << 1) | (unsigned)__double2loint(a)) != 0u) q = div_ieee(a, b);
    }
    return q;
}

/* the same split in two for straight-line code (gs_skew.cuh): the quotient, and a flag (0/1) that is set when the operand
 * lies outside the range where the fast sequence is exact; the caller ORs the flags of a group of points and redoes the
 * group with div_ieee in the (rare) flagged case */
__device__ __forceinline__ double div_inv_fast(const double a, const double b, const double y) {
    double q = a * y;
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-b, q, a);
    return __fma_rn(r, y, q);
}
__device__ __forceinline__ unsigned div_inv_bad(const double a) {
    const unsigned hi = (unsigned)__double2hiint(a);
    const unsigned e = (hi >> 20) & 0x7ffu;
    return (((e - 200u) > 1600u) && (((hi << 1) | (unsigned)__double2loint(a)) != 0u)) ? 1u : 0u;
}

/* ------------------------------------------------------------------------------------------------ */
/* point formulas                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

/* d3:140-147 / 156-163 / 201-208.  c?2 = h??*metric2, c?1 = h??*metric1 are the leading products of each
 * term (left-associated in the source, so forming them first is exact); den = 4.0*((cx2+cy2)+cz2). */
__device__ __forceinline__ double relax3(double pE, double pW, double pN, double pS, double pT, double pB, double rhs,
                                         double cx2, double cx1, double hx, double cy2, double cy1, double hy,
                                         double cz2, double cz1, double hz, double two_hxhyhz, double den) {
    double t = cx2 * (pE + pW) * 2.0;
    t = t + cx1 * (pE - pW) * hx;
    t = t + cy2 * (pN + pS) * 2.0;
    t = t + cy1 * (pN - pS) * hy;
    t = t + cz2 * (pT + pB) * 2.0;
    t = t + cz1 * (pT - pB) * hz;
    t = t - two_hxhyhz * rhs;
    return t / den;
}
/* the numerator of relax3 alone (same association) */
__device__ __forceinline__ double relax3_num(double pE, double pW, double pN, double pS, double pT, double pB, double rhs,
                                             double cx2, double cx1, double hx, double cy2, double cy1, double hy,
                                             double cz2, double cz1, double hz, double two_hxhyhz) {
    double t = cx2 * (pE + pW) * 2.0;
    t = t + cx1 * (pE - pW) * hx;
    t = t + cy2 * (pN + pS) * 2.0;
    t = t + cy1 * (pN - pS) * hy;
    t = t + cz2 * (pT + pB) * 2.0;
    t = t + cz1 * (pT - pB) * hz;
    t = t - two_hxhyhz * rhs;
    return t;
}

/* d2:121-126 / 133-138 / 160-165;  den = 4.0*(cx2 + cz2) */
__device__ __forceinline__ double relax2(double pE, double pW, double pT, double pB, double rhs,
                                         double cx2, double cx1, double hx, double cz2, double cz1, double hz,
                                         double two_hzhx, double den) {
    double t = cx2 * (pE + pW) * 2.0;
    t = t + cx1 * (pE - pW) * hx;
    t = t + cz2 * (pT + pB) * 2.0;
    t = t + cz1 * (pT - pB) * hz;
    t = t - two_hzhx * rhs;
    return t / den;
}
__device__ __forceinline__ double relax2_num(double pE, double pW, double pT, double pB, double rhs,
                                             double cx2, double cx1, double hx, double cz2, double cz1, double hz,
                                             double two_hzhx) {
    double t = cx2 * (pE + pW) * 2.0;
    t = t + cx1 * (pE - pW) * hx;
    t = t + cz2 * (pT + pB) * 2.0;
    t = t + cz1 * (pT - pB) * hz;
    t = t - two_hzhx * rhs;
    return t;
}

/* the Laplacian bracket of d3:100-105 / 219-224 / 411-416.  dxx = hx*hx, dx = 2.0*hx, ... */
__device__ __forceinline__ double lap3(double pC, double pE, double pW, double pN, double pS, double pT, double pB,
                                       double x2, double x1, double y2, double y1, double z2, double z1,
                                       double dxx, double dx, double dyy, double dy, double dzz, double dz) {
    double t = x2 * (pE - 2.0 * pC + pW) / dxx;
    t = t + x1 * (pE - pW) / dx;
    t = t + y2 * (pN - 2.0 * pC + pS) / dyy;
    t = t + y1 * (pN - pS) / dy;
    t = t + z2 * (pT - 2.0 * pC + pB) / dzz;
    t = t + z1 * (pT - pB) / dz;
    return t;
}

/* same bracket with the six divisors' reciprocals hoisted (div_inv): bit-identical to lap3 */
struct LapDiv { double dxx, dx, dyy, dy, dzz, dz, rxx, rx, ryy, ry, rzz, rz; };
__device__ __forceinline__ LapDiv make_lapdiv(double dxx, double dx, double dyy, double dy, double dzz, double dz, bool has_y) {
    LapDiv d;
    d.dxx = dxx; d.dx = dx; d.dyy = dyy; d.dy = dy; d.dzz = dzz; d.dz = dz;
    d.rxx = 1.0 / dxx; d.rx = 1.0 / dx; d.rzz = 1.0 / dzz; d.rz = 1.0 / dz;
    d.ryy = has_y ? 1.0 / dyy : 0.0; d.ry = has_y ? 1.0 / dy : 0.0;
    return d;
}
/* P2: every divisor is an exact power of two (always the case for SARAS grids: the computational coordinate runs over
 * [0,1] in 2^n intervals, grid.cc:155-164, so hx = 2^l * 2^-n).  Then a / b == a * (1/b) bit for bit -- the exact
 * quotient and the exact product are the same real number, rounded once -- and the six divisions cost six multiplies. */
template <bool P2>
__device__ __forceinline__ double div_by(const double a, const double b, const double y) { return P2 ? a * y : div_inv(a, b, y); }
template <bool P2>
__device__ __forceinline__ double lap3_inv(double pC, double pE, double pW, double pN, double pS, double pT, double pB,
                                           double x2, double x1, double y2, double y1, double z2, double z1, const LapDiv &d) {
    double t = div_by<P2>(x2 * (pE - 2.0 * pC + pW), d.dxx, d.rxx);
    t = t + div_by<P2>(x1 * (pE - pW), d.dx, d.rx);
    t = t + div_by<P2>(y2 * (pN - 2.0 * pC + pS), d.dyy, d.ryy);
    t = t + div_by<P2>(y1 * (pN - pS), d.dy, d.ry);
    t = t + div_by<P2>(z2 * (pT - 2.0 * pC + pB), d.dzz, d.rzz);
    t = t + div_by<P2>(z1 * (pT - pB), d.dz, d.rz);
    return t;
}
template <bool P2>
__device__ __forceinline__ double lap2_inv(double pC, double pE, double pW, double pT, double pB,
                                           double x2, double x1, double z2, double z1, const LapDiv &d) {
    double t = div_by<P2>(x2 * (pE - 2.0 * pC + pW), d.dxx, d.rxx);
    t = t + div_by<P2>(x1 * (pE - pW), d.dx, d.rx);
    t = t + div_by<P2>(z2 * (pT - 2.0 * pC + pB), d.dzz, d.rzz);
    t = t + div_by<P2>(z1 * (pT - pB), d.dz, d.rz);
    return t;
}

/* d2:98-101 / 174-177 (dxx = hx*hx) and d2:293-296 (dxx = hx2) */
__device__ __forceinline__ double lap2(double pC, double pE, double pW, double pT, double pB,
                                       double x2, double x1, double z2, double z1,
                                       double dxx, double dx, double dzz, double dz) {
    double t = x2 * (pE - 2.0 * pC + pW) / dxx;
    t = t + x1 * (pE - pW) / dx;
    t = t + z2 * (pT - 2.0 * pC + pB) / dzz;
    t = t + z1 * (pT - pB) / dz;
    return t;
}

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

/* ------------------------------------------------------------------------------------------------ */
/* Peer-memory halo PUSH of the Jacobi sweep (several ranks, CUDA IPC over NVLink).                     */
/* The sweep that computes a boundary plane also stores it into the neighbour's pad: every level array  */
/* of a rank lives in one slab that its face neighbours map, the slabs have the same layout on every     */
/* rank, so the neighbour's copy of the array being written is `its slab + the same offset`.  No pack,   */
/* no unpack, no exchange launch: the last CTA of the sweep raises a flag in each neighbour's arena, and  */
/* the face kernel that follows the sweep anyway (wall / periodic-z BCs of imposeBC) waits for the       */
/* neighbours' flags before it reads a pad.  Why this is safe with the Jacobi ping-pong (A -> B, B -> A): */
/* sweep n writes B's core here and B's pads at the neighbours, which nobody reads during sweep n; a      */
/* neighbour starts sweep n+1 (which writes A's pads here) only after its face kernel has seen this       */
/* rank's flag n, i.e. after this rank has finished reading A.  Only FACE pads are pushed (core rows of   */
/* the plane): the four edge columns (x pad and y pad at once) are not, see op_smooth in sarasmg.cu.      */
/* ------------------------------------------------------------------------------------------------ */
/* Bounded spin: a neighbour that never signals (a rank died, or ranks disagree on the exchange count) must not hang the
 * GPU.  After 2^27 polls (about a minute: ranks may legitimately be seconds apart) the waiter records the exchange
 * number in *err and goes on; the host turns a non-zero *err into an error at the next synchronisation point. */
__device__ __forceinline__ void ipc_wait_flag(const unsigned *flag, const unsigned epoch, unsigned *err) {
    if (!flag) return;
    if (err && *(volatile unsigned *)err) return;          /* an earlier wait already failed: do not stack timeouts */
    unsigned v;
    unsigned spins = 0;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - epoch) >= 0) return;
        if ((spins & 1023u) == 1023u) __nanosleep(200);
    } while (++spins < (1u << 27));
    if (err) atomicMax(err, epoch | 0x80000000u);
}
__device__ __forceinline__ void ipc_signal(unsigned *flag, const unsigned epoch) {
    if (flag) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
struct HaloPush {
    double *dst[4];             /* the neighbour's copy of the array this sweep writes: left, right, front, back (nullptr: none) */
    unsigned *sflag[4];         /* the flag to raise in that neighbour's arena */
    const unsigned *base;       /* exchange number = *base + add (see IpcPlan) */
    unsigned add;
    unsigned *counter;          /* arrival counter of this rank */
};
struct HaloWait {
    const unsigned *flag[4];    /* this rank's push flags, one per face (nullptr: nothing to wait for) */
    const unsigned *base;
    unsigned add;
    unsigned *err;
};
/* Every CTA of a pushing sweep calls this once, by all its threads, after its last store.  One thread fences for the CTA
 * (the barrier orders the other threads' stores before its fence, which is cumulative -- the grid.sync idiom): a
 * system-scope fence per warp costs microseconds, and CTAs that stored nothing remote need none. */
__device__ __forceinline__ void halo_push_signal(const HaloPush &H, const bool cta_pushed) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        if (cta_pushed) __threadfence_system();
        const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
        const unsigned t = atomicAdd(H.counter, 1u);
        if (t == nb - 1) {               /* the flag stores are release.sys: cumulative over what the atomics made visible here */
            *H.counter = 0u;
            const unsigned ep = *H.base + H.add;
#pragma unroll
            for (int f = 0; f < 4; f++) ipc_signal(H.sflag[f], ep);
        }
    }
}
__device__ __forceinline__ bool halo_cta_pushes(const HaloPush &H, const LevelView &L, const int i0, const int i1, const int j0, const int j1) {
    return (H.dst[0] && i0 <= 1 && 1 < i1) || (H.dst[1] && i0 <= L.nx - 2 && L.nx - 2 < i1) ||
           (H.dst[2] && j0 <= 1 && 1 < j1) || (H.dst[3] && j0 <= L.ny - 2 && L.ny - 2 < j1);
}
/* the (one or two) results of a thread's z pair at (i, j) go to the neighbours that own a pad there */
__device__ __forceinline__ void halo_push_store(const HaloPush &H, const LevelView &L, const int i, const int j, const int k0,
                                                const double2 v, const bool has2) {
    const long long zc = k0 + 1 + MG_LEAD;
#pragma unroll
    for (int f = 0; f < 4; f++) {
        if (!H.dst[f]) continue;
        long long o;
        if (f == 0) { if (i != 1) continue; o = (long long)(L.nx + 1) * L.plane + (long long)(j + 1) * L.pitch + zc; }
        else if (f == 1) { if (i != L.nx - 2) continue; o = (long long)(j + 1) * L.pitch + zc; }
        else if (f == 2) { if (j != 1) continue; o = (long long)(i + 1) * L.plane + (long long)(L.ny + 1) * L.pitch + zc; }
        else { if (j != L.ny - 2) continue; o = (long long)(i + 1) * L.plane + zc; }
        if (has2) *reinterpret_cast<double2 *>(H.dst[f] + o) = v; else H.dst[f][o] = v.x;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Jacobi sweep (d3:153-166, d2:131-140) and residual (d3:96-108, d2:95-103): x-marching kernels      */
/* Thread = one aligned z pair (k0, k0+1) of row j; marches i over [i0, i1).                          */
/* Algorithmic traffic: 24 B per core point (read p, read rhs, write out).                            */
/* ------------------------------------------------------------------------------------------------ */
#ifndef MG_MARCH_MINB
#define MG_MARCH_MINB 4      /* resident CTAs per SM the marching kernels are compiled for (register cap 65536 / (256 * MINB)) */
#endif
#ifndef MG_ERROR_MINB
#define MG_ERROR_MINB 4
#endif
/* software prefetch distance of the marching kernels, in x planes (0 = off): each thread asks L2 for the lines it will
 * load MG_PF_DIST iterations later, so the demand loads see L2 latency instead of HBM latency without holding registers */
#ifndef MG_PF_DIST
#define MG_PF_DIST 0
#endif
#ifndef MG_PROLONG_PF
#define MG_PROLONG_PF 0
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int DIM, int MODE /* 0 Jacobi, 1 residual */, bool XU /* MODE 0: xix2 constant on this level, the Jacobi diagonal is march-invariant;
                                                                 MODE 1: the six divisors are powers of two (div_by) */, bool PUSH>
__device__ __forceinline__ void stencil_march_body(const double *__restrict__ P, const double *__restrict__ R, double *__restrict__ O,
                                                   const LevelView &L, const int ichunk, const HaloPush *H) {
    const int k0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int j = (DIM == 3) ? (blockIdx.y * blockDim.y + threadIdx.y) : 0;
    if (k0 >= L.nz || j >= L.ny) return;
    const int i0 = blockIdx.z * ichunk;
    const int i1 = min(i0 + ichunk, L.nx);
    const bool has2 = (k0 + 1 < L.nz);
    const long long plane = L.plane;
    const int pitch = L.pitch;
    const long long row = (long long)(j + 1) * pitch + (k0 + 1 + MG_LEAD);
    const double *p = P + row;
    const double *r = R + row;
    double *o = O + row;

    /* per-thread constants */
    const double z2a = L.ztz2[k0], z2b = L.ztz2[k0 + 1], z1a = L.ztzz[k0], z1b = L.ztzz[k0 + 1];
    double y2 = 0.0, y1 = 0.0;
    if (DIM == 3) { y2 = L.ety2[j]; y1 = L.etyy[j]; }
    const double hx = L.hx, hy = L.hy, hz = L.hz;

    /* MODE 0 */
    double cy2 = 0.0, cy1 = 0.0, cz2a, cz2b, cz1a, cz1b, two_h, kx;
    if (DIM == 3) {
        cy2 = L.hzhx * y2; cy1 = L.hzhx * y1;
        cz2a = L.hxhy * z2a; cz2b = L.hxhy * z2b; cz1a = L.hxhy * z1a; cz1b = L.hxhy * z1b;
        two_h = 2.0 * L.hxhyhz; kx = L.hyhz;
    } else {
        cz2a = L.hx2 * z2a; cz2b = L.hx2 * z2b; cz1a = L.hx2 * z1a; cz1b = L.hx2 * z1b;
        two_h = 2.0 * L.hzhx; kx = L.hz2;
    }
    /* MODE 1: the six divisors hx*hx, 2.0*hx, ... are march-invariant */
    LapDiv dv;
    if (MODE == 1) dv = make_lapdiv(hx * hx, 2.0 * hx, hy * hy, 2.0 * hy, hz * hz, 2.0 * hz, DIM == 3);
    /* MODE 0 with XU: den = 4.0*((cx2+cy2)+cz2) is march-invariant too */
    double den_a = 1.0, den_b = 1.0, rden_a = 1.0, rden_b = 1.0;
    if (MODE == 0 && XU) {
        const double cx2 = kx * L.xix2[i0];
        if (DIM == 3) { const double s = cx2 + cy2; den_a = 4.0 * (s + cz2a); den_b = 4.0 * (s + cz2b); }
        else { den_a = 4.0 * (cx2 + cz2a); den_b = 4.0 * (cx2 + cz2b); }
        rden_a = 1.0 / den_a; rden_b = 1.0 / den_b;
    }

    const bool push_row = PUSH && ((H->dst[2] && j == 1) || (H->dst[3] && j == L.ny - 2));
    double2 pW = ld2(p + (long long)i0 * plane);          /* plane i0-1 */
    double2 pC = ld2(p + (long long)(i0 + 1) * plane);    /* plane i0   */
    for (int i = i0; i < i1; ++i) {
        const long long b = (long long)(i + 1) * plane;
        if (MG_PF_DIST > 0 && i + MG_PF_DIST < i1) {
            prefetch_l2(p + b + (long long)(MG_PF_DIST + 1) * plane);
            prefetch_l2(r + b + (long long)MG_PF_DIST * plane);
        }
        const double2 pE = ld2(p + b + plane);
        const double2 rr = ld2(r + b);
        const double pBa = p[b - 1];        /* (i, j, k0-1) */
        const double pTb = p[b + 2];        /* (i, j, k0+2) */
        double2 pS = make_double2(0.0, 0.0), pN = make_double2(0.0, 0.0);
        if (DIM == 3) { pS = ld2(p + b - pitch); pN = ld2(p + b + pitch); }
        const double x2 = L.xix2[i], x1 = L.xixx[i];
        double2 out;
        if (MODE == 0) {
            const double cx2 = kx * x2, cx1 = kx * x1;
            if (XU) {
                if (DIM == 3) {
                    out.x = div_inv(relax3_num(pE.x, pW.x, pN.x, pS.x, pC.y, pBa, rr.x, cx2, cx1, hx, cy2, cy1, hy, cz2a, cz1a, hz, two_h), den_a, rden_a);
                    out.y = div_inv(relax3_num(pE.y, pW.y, pN.y, pS.y, pTb, pC.x, rr.y, cx2, cx1, hx, cy2, cy1, hy, cz2b, cz1b, hz, two_h), den_b, rden_b);
                } else {
                    out.x = div_inv(relax2_num(pE.x, pW.x, pC.y, pBa, rr.x, cx2, cx1, hx, cz2a, cz1a, hz, two_h), den_a, rden_a);
                    out.y = div_inv(relax2_num(pE.y, pW.y, pTb, pC.x, rr.y, cx2, cx1, hx, cz2b, cz1b, hz, two_h), den_b, rden_b);
                }
            } else if (DIM == 3) {
                const double s = cx2 + cy2;
                out.x = relax3(pE.x, pW.x, pN.x, pS.x, pC.y, pBa, rr.x, cx2, cx1, hx, cy2, cy1, hy, cz2a, cz1a, hz, two_h, 4.0 * (s + cz2a));
                out.y = relax3(pE.y, pW.y, pN.y, pS.y, pTb, pC.x, rr.y, cx2, cx1, hx, cy2, cy1, hy, cz2b, cz1b, hz, two_h, 4.0 * (s + cz2b));
            } else {
                out.x = relax2(pE.x, pW.x, pC.y, pBa, rr.x, cx2, cx1, hx, cz2a, cz1a, hz, two_h, 4.0 * (cx2 + cz2a));
                out.y = relax2(pE.y, pW.y, pTb, pC.x, rr.y, cx2, cx1, hx, cz2b, cz1b, hz, two_h, 4.0 * (cx2 + cz2b));
            }
        } else {
            if (DIM == 3) {
                out.x = rr.x - lap3_inv<XU>(pC.x, pE.x, pW.x, pN.x, pS.x, pC.y, pBa, x2, x1, y2, y1, z2a, z1a, dv);
                out.y = rr.y - lap3_inv<XU>(pC.y, pE.y, pW.y, pN.y, pS.y, pTb, pC.x, x2, x1, y2, y1, z2b, z1b, dv);
            } else {
                out.x = rr.x - lap2_inv<XU>(pC.x, pE.x, pW.x, pC.y, pBa, x2, x1, z2a, z1a, dv);
                out.y = rr.y - lap2_inv<XU>(pC.y, pE.y, pW.y, pTb, pC.x, x2, x1, z2b, z1b, dv);
            }
        }
        if (has2) st2(o + b, out); else o[b] = out.x;
        if (PUSH && (push_row | (i == 1) | (i == L.nx - 2))) halo_push_store(*H, L, i, j, k0, out, has2);
        pW = pC; pC = pE;
    }
}
template <int DIM, int MODE, bool XU>
__global__ void __launch_bounds__(256, MG_MARCH_MINB)
k_stencil_march(const double *__restrict__ P, const double *__restrict__ R, double *__restrict__ O, const LevelView L,
                const int ichunk) {
    stencil_march_body<DIM, MODE, XU, false>(P, R, O, L, ichunk, nullptr);
}
/* Jacobi sweep of a rank with neighbours: the same arithmetic, plus the halo push (3-D only) */
template <bool XU>
__global__ void __launch_bounds__(256, MG_MARCH_MINB)
k_stencil_march_push(const double *__restrict__ P, const double *__restrict__ R, double *__restrict__ O, const LevelView L,
                     const int ichunk, const HaloPush H) {
    stencil_march_body<3, 0, XU, true>(P, R, O, L, ichunk, &H);
    const int i0 = blockIdx.z * ichunk, j0 = blockIdx.y * blockDim.y;
    halo_push_signal(H, halo_cta_pushes(H, L, i0, min(i0 + ichunk, L.nx), j0, j0 + (int)blockDim.y));
}

/* ------------------------------------------------------------------------------------------------ */
/* computeError (d3:408-430, d2:291-309): |L p - rhs| reduced to max / sum of squares; no store.       */
/* partial layout: part[0*nb + b] = num, part[1*nb + b] = den(sum rhs^2), b = linear block id.         */
/* ------------------------------------------------------------------------------------------------ */
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
/* deterministic block reduction (fixed shuffle tree); result valid in thread 0 */
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v, double *sm /* >= 32 doubles */) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthr = blockDim.x * blockDim.y;
    const int lane = tid & 31, w = tid >> 5;
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    const int nw = (nthr + 31) >> 5;
    if (w == 0) {
        v = (lane < nw) ? sm[lane] : (IS_MAX ? -1.0e300 : 0.0);
        v = IS_MAX ? warp_max(v) : warp_sum(v);
    }
    return v;
}

/* Compensated (cascaded) summation for the RHS mean of mgSolve (p:122): the running sum s carries the rounding error of
 * every addition in c (Knuth two-sum; exact under -fmad=false), so s + c is the exact sum to ~2^-100 and its rounding is the
 * correctly rounded sum whatever the reduction tree.  The reference's own value is ONE sequential fp64 sum whose rounding
 * (~sqrt(n) ulp) no parallel order reproduces; the device value is the exact mean instead. */
struct csum { double s, c; };
__device__ __forceinline__ void cs_add(csum &a, const double x) {
    const double t = a.s + x;
    const double bb = t - a.s;
    a.c += (a.s - (t - bb)) + (x - bb);
    a.s = t;
}
__device__ __forceinline__ void cs_merge(csum &a, const double s2, const double c2) {
    cs_add(a, s2);
    a.c += c2;
}
/* deterministic block reduction of a compensated sum; result valid in thread 0; sm >= 64 doubles */
__device__ __forceinline__ csum block_reduce_cs(csum v, double *sm) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthr = blockDim.x * blockDim.y;
    const int lane = tid & 31, w = tid >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        const double s2 = __shfl_down_sync(0xffffffffu, v.s, o), c2 = __shfl_down_sync(0xffffffffu, v.c, o);
        cs_merge(v, s2, c2);
    }
    __syncthreads();
    if (lane == 0) { sm[2 * w] = v.s; sm[2 * w + 1] = v.c; }
    __syncthreads();
    const int nw = (nthr + 31) >> 5;
    if (w == 0) {
        v.s = (lane < nw) ? sm[2 * lane] : 0.0;
        v.c = (lane < nw) ? sm[2 * lane + 1] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            const double s2 = __shfl_down_sync(0xffffffffu, v.s, o), c2 = __shfl_down_sync(0xffffffffu, v.c, o);
            cs_merge(v, s2, c2);
        }
    }
    return v;
}
/* final stage for compensated partials: part[2b], part[2b+1] = (s, c) of block b; out[0] = fl(s + c) of the total */
__global__ void __launch_bounds__(1024) k_reduce_final_cs(const double *__restrict__ part, const int n, double *__restrict__ out) {
    __shared__ double sm[64];
    csum v = {0.0, 0.0};
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) cs_merge(v, part[2 * idx], part[2 * idx + 1]);
    v = block_reduce_cs(v, sm);
    if (threadIdx.x == 0) out[0] = v.s + v.c;
}

template <int DIM, bool P2>
__global__ void __launch_bounds__(256, MG_ERROR_MINB)
k_error_march(const double *__restrict__ P, const double *__restrict__ R, const LevelView L, const int ichunk,
              const int norm_order, double hx_c, double hy_c, double hz_c, double *__restrict__ part, const int nblocks) {
    __shared__ double sm[32];
    const int k0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int j = (DIM == 3) ? (blockIdx.y * blockDim.y + threadIdx.y) : 0;
    const bool active = !(k0 >= L.nz || j >= L.ny);
    double num = 0.0, den = 0.0;
    if (active) {
        const int i0 = blockIdx.z * ichunk;
        const int i1 = min(i0 + ichunk, L.nx);
        const bool has2 = (k0 + 1 < L.nz);
        const long long plane = L.plane;
        const int pitch = L.pitch;
        const long long row = (long long)(j + 1) * pitch + (k0 + 1 + MG_LEAD);
        const double *p = P + row;
        const double *r = R + row;
        const double z2a = L.ztz2[k0], z2b = L.ztz2[k0 + 1], z1a = L.ztzz[k0], z1b = L.ztzz[k0 + 1];
        double y2 = 0.0, y1 = 0.0;
        if (DIM == 3) { y2 = L.ety2[j]; y1 = L.etyy[j]; }
        /* d3:411-416 divides by hx(vLevel)*hx(vLevel); d2:293-296 by hx2(0), hz2(0) */
        const double dxx = (DIM == 3) ? hx_c * hx_c : L.hx2, dx = 2.0 * hx_c;
        const double dyy = hy_c * hy_c, dy = 2.0 * hy_c;
        const double dzz = (DIM == 3) ? hz_c * hz_c : L.hz2, dz = 2.0 * hz_c;
        const LapDiv dv = make_lapdiv(dxx, dx, dyy, dy, dzz, dz, DIM == 3);
        double2 pW = ld2(p + (long long)i0 * plane);
        double2 pC = ld2(p + (long long)(i0 + 1) * plane);
        for (int i = i0; i < i1; ++i) {
            const long long b = (long long)(i + 1) * plane;
            if (MG_PF_DIST > 0 && i + MG_PF_DIST < i1) {
                prefetch_l2(p + b + (long long)(MG_PF_DIST + 1) * plane);
                prefetch_l2(r + b + (long long)MG_PF_DIST * plane);
            }
            const double2 pE = ld2(p + b + plane);
            const double2 rr = ld2(r + b);
            const double pBa = p[b - 1];
            const double pTb = p[b + 2];
            double2 pS = make_double2(0.0, 0.0), pN = make_double2(0.0, 0.0);
            if (DIM == 3) { pS = ld2(p + b - pitch); pN = ld2(p + b + pitch); }
            const double x2 = L.xix2[i], x1 = L.xixx[i];
            double ta, tb;
            if (DIM == 3) {
                ta = fabs(lap3_inv<P2>(pC.x, pE.x, pW.x, pN.x, pS.x, pC.y, pBa, x2, x1, y2, y1, z2a, z1a, dv) - rr.x);
                tb = fabs(lap3_inv<P2>(pC.y, pE.y, pW.y, pN.y, pS.y, pTb, pC.x, x2, x1, y2, y1, z2b, z1b, dv) - rr.y);
            } else {
                ta = fabs(lap2_inv<P2>(pC.x, pE.x, pW.x, pC.y, pBa, x2, x1, z2a, z1a, dv) - rr.x);
                tb = fabs(lap2_inv<P2>(pC.y, pE.y, pW.y, pTb, pC.x, x2, x1, z2b, z1b, dv) - rr.y);
            }
            if (norm_order == 1) {
                num = fmax(num, ta);
                if (has2) num = fmax(num, tb);
            } else {
                num += ta * ta; den += rr.x * rr.x;
                if (has2) { num += tb * tb; den += rr.y * rr.y; }
            }
            pW = pC; pC = pE;
        }
    }
    const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    double n2 = (norm_order == 1) ? block_reduce<true>(num, sm) : block_reduce<false>(num, sm);
    double d2 = block_reduce<false>(den, sm);
    if (threadIdx.x == 0 && threadIdx.y == 0) { part[bid] = n2; part[nblocks + bid] = d2; }
}

/* final stage: one block reduces `n` partials of `nq` quantities in a fixed order.
 * is_max bit q set -> quantity q is a max-reduction.  out[q] receives the result. */
__global__ void __launch_bounds__(1024) k_reduce_final(const double *__restrict__ part, const int n, const int nq,
                                                      const unsigned is_max, double *__restrict__ out) {
    __shared__ double sm[32];
    for (int q = 0; q < nq; ++q) {
        const bool mx = (is_max >> q) & 1u;
        double v = mx ? -1.0e300 : 0.0;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const double x = part[(long long)q * n + idx];
            v = mx ? fmax(v, x) : v + x;
        }
        v = mx ? block_reduce<true>(v, sm) : block_reduce<false>(v, sm);
        if (threadIdx.x == 0) out[q] = v;
        __syncthreads();
    }
}

/* reductions over the WHOLE padded array (blitz::mean p:122; blitz::max(fabs()) d3:437): one block per (i,j) row group */
template <int WHAT /* 0 sum, 1 absmax */>
__global__ void __launch_bounds__(256) k_reduce_padded(const double *__restrict__ A, const LevelView L, double *__restrict__ part) {
    __shared__ double sm[32];
    const long long rows = (long long)L.Fx * L.Fy;
    double v = (WHAT == 1) ? -1.0e300 : 0.0;
    for (long long rrow = blockIdx.x; rrow < rows; rrow += gridDim.x) {
        const double *a = A + rrow * L.pitch + MG_LEAD;
        for (int k = threadIdx.x; k < L.Fz; k += blockDim.x) {
            const double x = a[k];
            v = (WHAT == 1) ? fmax(v, fabs(x)) : v + x;
        }
    }
    v = (WHAT == 1) ? block_reduce<true>(v, sm) : block_reduce<false>(v, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}

/* residualData(0) -= globalAvg over the whole padded array (p:128) */
__global__ void __launch_bounds__(256) k_sub_scalar_padded(double *__restrict__ A, const LevelView L, const double v) {
    const long long rows = (long long)L.Fx * L.Fy;
    for (long long rrow = blockIdx.x; rrow < rows; rrow += gridDim.x) {
        double *a = A + rrow * L.pitch + MG_LEAD;
        for (int k = threadIdx.x; k < L.Fz; k += blockDim.x) a[k] = a[k] - v;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* mgSolve preamble (p:107-128) in three streaming passes instead of memset x2 + copy x2 + reduce + sub:  */
/*   k_load_inputs   pressureData(0) / residualData(0) = 0 everywhere, core = core of inFn / rhs (p:107-115),   */
/*                   and the sum of the rhs core (first part of blitz::mean, p:122), 32 B per point;       */
/*   k_reduce_shell  sum of the pads of residualData(0) after updatePads (the rest of that mean);          */
/*   k_sub_rows      residualData(0) -= globalAvg over the whole padded array (p:128), 16 B per point.      */
/* Rows are walked with aligned 128-bit accesses: column c of a row holds k = c - (MG_LEAD + 1).            */
/* ------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_load_inputs(const double *__restrict__ inP, const double *__restrict__ inR,
                                                     double *__restrict__ P, double *__restrict__ R, const LevelView L,
                                                     double *__restrict__ part) {
    __shared__ double sm[64];
    const long long rows = (long long)L.Fx * L.Fy;
    const int npair = L.pitch / 2;
    csum acc = {0.0, 0.0};
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int i = (int)(row / L.Fy) - 1, j = (int)(row % L.Fy) - 1;
        const bool core_row = (i >= 0 && i < L.nx && j >= 0 && j < L.ny);
        const long long base = row * L.pitch;
        for (int q = threadIdx.x; q < npair; q += blockDim.x) {
            const int k = 2 * q - (MG_LEAD + 1);
            double2 p = make_double2(0.0, 0.0), r = make_double2(0.0, 0.0);
            if (core_row && k + 1 >= 0 && k < L.nz) {
                const double2 a = ld2(inP + base + 2 * q), b = ld2(inR + base + 2 * q);
                if (k >= 0) { p.x = a.x; r.x = b.x; cs_add(acc, b.x); }
                if (k + 1 < L.nz) { p.y = a.y; r.y = b.y; cs_add(acc, b.y); }
            }
            st2(P + base + 2 * q, p);
            st2(R + base + 2 * q, r);
        }
    }
    acc = block_reduce_cs(acc, sm);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = acc.s; part[2 * blockIdx.x + 1] = acc.c; }
}

/* sum over the pad shell of a level array: x-pad planes, y-pad rows of the core planes, z-pad ends of the core rows */
__global__ void __launch_bounds__(256) k_reduce_shell(const double *__restrict__ A, const LevelView L, double *__restrict__ part) {
    __shared__ double sm[64];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const long long nXp = 2LL * L.Fy * L.Fz;                    /* planes i = -1, nx: every (j, k) */
    const long long nYp = 2LL * L.nx * L.Fz;                    /* rows j = -1, ny of the core planes: every k */
    const long long nZp = 2LL * L.nx * L.ny;                    /* k = -1, nz of the core rows */
    csum acc = {0.0, 0.0};
    for (long long q = tid; q < nXp + nYp + nZp; q += nthr) {
        int i, j, k;
        if (q < nXp) {
            const long long w = q % ((long long)L.Fy * L.Fz);
            i = (q < (long long)L.Fy * L.Fz) ? -1 : L.nx; j = (int)(w / L.Fz) - 1; k = (int)(w % L.Fz) - 1;
        } else if (q < nXp + nYp) {
            const long long w = q - nXp, h = (long long)L.nx * L.Fz;
            j = (w < h) ? -1 : L.ny; i = (int)((w % h) / L.Fz); k = (int)(w % L.Fz) - 1;
        } else {
            const long long w = q - nXp - nYp, h = (long long)L.nx * L.ny;
            k = (w < h) ? -1 : L.nz; i = (int)((w % h) / L.ny); j = (int)(w % L.ny);
        }
        cs_add(acc, A[mg_at(L, i, j, k)]);
    }
    acc = block_reduce_cs(acc, sm);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = acc.s; part[2 * blockIdx.x + 1] = acc.c; }
}

__global__ void __launch_bounds__(256) k_sub_rows(double *__restrict__ A, const LevelView L, const double v) {
    const long long rows = (long long)L.Fx * L.Fy;
    const int npair = L.pitch / 2;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const long long base = row * L.pitch;
        for (int q = threadIdx.x; q < npair; q += blockDim.x) {
            const int k = 2 * q - (MG_LEAD + 1);                /* the padded array covers k = -1 .. nz */
            if (k + 1 < -1 || k > L.nz) continue;
            double2 a = ld2(A + base + 2 * q);
            if (k >= -1) a.x = a.x - v;
            if (k + 1 <= L.nz) a.y = a.y - v;
            st2(A + base + 2 * q, a);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* face operations: physical BCs (imposeBC d3:578-667, d2:406-463), periodic self-neighbour halos      */
/* (updatePads with faceRanks == own rank, d3:675-685) and pad zeroing (d3:91).  Strided-loop form so   */
/* the same code runs grid-wide and inside the single-CTA coarse solver.                               */
/* ------------------------------------------------------------------------------------------------ */
template <int DIR>
__device__ __forceinline__ void face_op(double *A, const LevelView &L, const int lo_mode, const int hi_mode,
                                        const double *wall, const bool planar, const long long tid, const long long nthr) {
    const int jlo = planar ? 0 : -1;
    const int nj = planar ? 1 : L.Fy;
    if (DIR == 0) {
        const int xE = L.nx - 1;
        const long long cnt = (long long)nj * L.Fz;
        for (long long q = tid; q < cnt; q += nthr) {
            const int j = jlo + (int)(q / L.Fz), k = (int)(q % L.Fz) - 1;
            const double w = (lo_mode == FACE_WALL || hi_mode == FACE_WALL) ? (planar ? wall[k + 1] : wall[(long long)(j + 1) * L.Fz + (k + 1)]) : 0.0;
            if (lo_mode != FACE_NONE) {
                double v;
                if (lo_mode == FACE_MIRROR) v = A[mg_at(L, 1, j, k)];
                else if (lo_mode == FACE_NEG_MIRROR) v = -A[mg_at(L, 1, j, k)];
                else if (lo_mode == FACE_WALL) v = w;
                else if (lo_mode == FACE_PERIODIC) v = A[mg_at(L, xE - 1, j, k)];
                else v = 0.0;
                A[mg_at(L, -1, j, k)] = v;
            }
            if (hi_mode != FACE_NONE) {
                double v;
                if (hi_mode == FACE_MIRROR) v = A[mg_at(L, xE - 1, j, k)];
                else if (hi_mode == FACE_NEG_MIRROR) v = -A[mg_at(L, xE - 1, j, k)];
                else if (hi_mode == FACE_WALL) v = w;
                else if (hi_mode == FACE_PERIODIC) v = A[mg_at(L, 1, j, k)];
                else v = 0.0;
                A[mg_at(L, xE + 1, j, k)] = v;
            }
        }
    } else if (DIR == 1) {
        const int yE = L.ny - 1;
        const long long cnt = (long long)L.Fx * L.Fz;
        for (long long q = tid; q < cnt; q += nthr) {
            const int i = (int)(q / L.Fz) - 1, k = (int)(q % L.Fz) - 1;
            const double w = (lo_mode == FACE_WALL || hi_mode == FACE_WALL) ? wall[(long long)(i + 1) * L.Fz + (k + 1)] : 0.0;
            if (lo_mode != FACE_NONE) {
                double v;
                if (lo_mode == FACE_MIRROR) v = A[mg_at(L, i, 1, k)];
                else if (lo_mode == FACE_NEG_MIRROR) v = -A[mg_at(L, i, 1, k)];
                else if (lo_mode == FACE_WALL) v = w;
                else if (lo_mode == FACE_PERIODIC) v = A[mg_at(L, i, yE - 1, k)];
                else v = 0.0;
                A[mg_at(L, i, -1, k)] = v;
            }
            if (hi_mode != FACE_NONE) {
                double v;
                if (hi_mode == FACE_MIRROR) v = A[mg_at(L, i, yE - 1, k)];
                else if (hi_mode == FACE_NEG_MIRROR) v = -A[mg_at(L, i, yE - 1, k)];
                else if (hi_mode == FACE_WALL) v = w;
                else if (hi_mode == FACE_PERIODIC) v = A[mg_at(L, i, 1, k)];
                else v = 0.0;
                A[mg_at(L, i, yE + 1, k)] = v;
            }
        }
    } else {
        const int zE = L.nz - 1;
        const long long cnt = (long long)L.Fx * nj;
        for (long long q = tid; q < cnt; q += nthr) {
            const int i = (int)(q / nj) - 1, j = jlo + (int)(q % nj);
            const double w = (lo_mode == FACE_WALL || hi_mode == FACE_WALL) ? (planar ? wall[i + 1] : wall[(long long)(i + 1) * L.Fy + (j + 1)]) : 0.0;
            if (lo_mode != FACE_NONE) {
                double v;
                if (lo_mode == FACE_MIRROR) v = A[mg_at(L, i, j, 1)];
                else if (lo_mode == FACE_NEG_MIRROR) v = -A[mg_at(L, i, j, 1)];
                else if (lo_mode == FACE_WALL) v = w;
                else if (lo_mode == FACE_PERIODIC) v = A[mg_at(L, i, j, zE - 1)];
                else v = 0.0;
                A[mg_at(L, i, j, -1)] = v;
            }
            if (hi_mode != FACE_NONE) {
                double v;
                if (hi_mode == FACE_MIRROR) v = A[mg_at(L, i, j, zE - 1)];
                else if (hi_mode == FACE_NEG_MIRROR) v = -A[mg_at(L, i, j, zE - 1)];
                else if (hi_mode == FACE_WALL) v = w;
                else if (hi_mode == FACE_PERIODIC) v = A[mg_at(L, i, j, 1)];
                else v = 0.0;
                A[mg_at(L, i, j, zE + 1)] = v;
            }
        }
    }
}

template <int DIR>
__global__ void __launch_bounds__(256) k_face(double *A, const LevelView L, const int lo_mode, const int hi_mode,
                                              const double *wall, const int planar) {
    face_op<DIR>(A, L, lo_mode, hi_mode, wall, planar != 0, (long long)blockIdx.x * blockDim.x + threadIdx.x,
                 (long long)gridDim.x * blockDim.x);
}

/* ------------------------------------------------------------------------------------------------ */
/* All six face operations of one imposeBC / self-neighbour updatePads / pad zeroing in ONE launch.     */
/* The reference applies x, then y, then z, each over full planes including the pads of the other        */
/* directions (d3:578-667), so an edge or corner pad ends up with the value of the LAST operation that    */
/* writes it, and that operation read a pad the previous one had written.  Here every pad point is        */
/* written once, by a thread of its last-writing operation, which resolves the chain of sources           */
/* backwards (z source, then y, then x) down to a point this kernel does not write (a core point, or a   */
/* pad whose direction has no operation: halo data).  Same values, no ordering between threads needed.   */
/* ------------------------------------------------------------------------------------------------ */
struct FaceModes { int m[6]; const double *xwall, *ywall, *zwall; };      /* x lo, x hi, y lo, y hi, z lo, z hi */

__device__ __forceinline__ int face_src(const int mode, const bool lo, const int E) {   /* source index of a pad along one axis */
    if (mode == FACE_PERIODIC) return lo ? E - 1 : 1;
    return lo ? 1 : E - 1;                              /* mirror, negative mirror */
}
template <int DIM>
__device__ __forceinline__ double face_resolve(const double *A, const LevelView &L, const FaceModes &F, int i, int j, int k) {
    const int xE = L.nx - 1, yE = L.ny - 1, zE = L.nz - 1;
    bool neg = false;
    {
        const int side = (k == -1) ? 4 : (k == zE + 1) ? 5 : -1;
        const int md = side >= 0 ? F.m[side] : FACE_NONE;
        if (md != FACE_NONE) {
            if (md == FACE_ZERO) return 0.0;
            if (md == FACE_WALL) return (DIM == 2) ? F.zwall[i + 1] : F.zwall[(long long)(i + 1) * L.Fy + (j + 1)];
            k = face_src(md, side == 4, zE);
            neg ^= (md == FACE_NEG_MIRROR);
        }
    }
    if (DIM == 3) {
        const int side = (j == -1) ? 2 : (j == yE + 1) ? 3 : -1;
        const int md = side >= 0 ? F.m[side] : FACE_NONE;
        if (md != FACE_NONE) {
            if (md == FACE_ZERO) return 0.0;
            if (md == FACE_WALL) { const double w = F.ywall[(long long)(i + 1) * L.Fz + (k + 1)]; return neg ? -w : w; }
            j = face_src(md, side == 2, yE);
            neg ^= (md == FACE_NEG_MIRROR);
        }
    }
    {
        const int side = (i == -1) ? 0 : (i == xE + 1) ? 1 : -1;
        const int md = side >= 0 ? F.m[side] : FACE_NONE;
        if (md != FACE_NONE) {
            if (md == FACE_ZERO) return 0.0;
            if (md == FACE_WALL) { const double w = (DIM == 2) ? F.xwall[k + 1] : F.xwall[(long long)(j + 1) * L.Fz + (k + 1)]; return neg ? -w : w; }
            i = face_src(md, side == 0, xE);
            neg ^= (md == FACE_NEG_MIRROR);
        }
    }
    const double v = A[mg_at(L, i, j, k)];
    return neg ? -v : v;
}

template <int DIM>
__device__ __forceinline__ void faces_apply(double *A, const LevelView &L, const FaceModes &F, const long long tid, const long long nthr) {
    const int jlo = (DIM == 2) ? 0 : -1, nj = (DIM == 2) ? 1 : L.Fy;
    const int yE = L.ny - 1, zE = L.nz - 1, xE = L.nx - 1;
    const long long nZ = (F.m[4] | F.m[5]) ? (long long)L.Fx * nj : 0;                   /* (i, j) columns of the z faces */
    const long long nY = (DIM == 3 && (F.m[2] | F.m[3])) ? (long long)L.Fx * L.Fz : 0;   /* (i, k) of the y faces */
    const long long nX = (F.m[0] | F.m[1]) ? (long long)nj * L.Fz : 0;                   /* (j, k) of the x faces */
    const long long tot = nZ + nY + nX;
    for (long long q = tid; q < tot; q += nthr) {
        if (q < nZ) {
            const int i = (int)(q / nj) - 1, j = jlo + (int)(q % nj);
            if (F.m[4] != FACE_NONE) A[mg_at(L, i, j, -1)] = face_resolve<DIM>(A, L, F, i, j, -1);
            if (F.m[5] != FACE_NONE) A[mg_at(L, i, j, zE + 1)] = face_resolve<DIM>(A, L, F, i, j, zE + 1);
        } else if (q < nZ + nY) {
            const long long w = q - nZ;
            const int i = (int)(w / L.Fz) - 1, k = (int)(w % L.Fz) - 1;
            /* a z operation on this side writes the point after the y operation: its thread owns it */
            if ((k == -1 && F.m[4] != FACE_NONE) || (k == zE + 1 && F.m[5] != FACE_NONE)) continue;
            if (F.m[2] != FACE_NONE) A[mg_at(L, i, -1, k)] = face_resolve<DIM>(A, L, F, i, -1, k);
            if (F.m[3] != FACE_NONE) A[mg_at(L, i, yE + 1, k)] = face_resolve<DIM>(A, L, F, i, yE + 1, k);
        } else {
            const long long w = q - nZ - nY;
            const int j = jlo + (int)(w / L.Fz), k = (int)(w % L.Fz) - 1;
            if ((k == -1 && F.m[4] != FACE_NONE) || (k == zE + 1 && F.m[5] != FACE_NONE)) continue;
            if (DIM == 3 && ((j == -1 && F.m[2] != FACE_NONE) || (j == yE + 1 && F.m[3] != FACE_NONE))) continue;
            if (F.m[0] != FACE_NONE) A[mg_at(L, -1, j, k)] = face_resolve<DIM>(A, L, F, -1, j, k);
            if (F.m[1] != FACE_NONE) A[mg_at(L, xE + 1, j, k)] = face_resolve<DIM>(A, L, F, xE + 1, j, k);
        }
    }
}
template <int DIM>
__global__ void __launch_bounds__(256) k_faces(double *A, const LevelView L, const FaceModes F) {
    faces_apply<DIM>(A, L, F, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}
/* the face kernel after a pushing Jacobi sweep: wait for the neighbours' pushes (HaloPush), then the same face operations */
template <int DIM>
__global__ void __launch_bounds__(256) k_faces_wait(double *A, const LevelView L, const FaceModes F, const HaloWait W) {
    if (threadIdx.x == 0) {
        const unsigned ep = *W.base + W.add;
#pragma unroll
        for (int f = 0; f < 4; f++) ipc_wait_flag(W.flag[f], ep, W.err);
    }
    __syncthreads();
    faces_apply<DIM>(A, L, F, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

/* y-row pack/unpack for the inter-GPU halo (MPI_Type_vector of d3:492-497): row j of every x plane <-> contiguous buffer */
__global__ void __launch_bounds__(256) k_pack_y(const double *__restrict__ A, const LevelView L, const int j, double *__restrict__ buf) {
    const long long cnt = (long long)L.Fx * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / L.Fz) - 1, k = (int)(q % L.Fz) - 1;
        buf[q] = A[mg_at(L, i, j, k)];
    }
}
__global__ void __launch_bounds__(256) k_unpack_y(double *__restrict__ A, const LevelView L, const int j, const double *__restrict__ buf) {
    const long long cnt = (long long)L.Fx * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / L.Fz) - 1, k = (int)(q % L.Fz) - 1;
        A[mg_at(L, i, j, k)] = buf[q];
    }
}
/* x-plane pack/unpack (MPI_Type_contiguous of d3:486-489; 2-D: row j=0 only, d2:353-356) into a compact buffer */
__global__ void __launch_bounds__(256) k_pack_x(const double *__restrict__ A, const LevelView L, const int i, const int planar, double *__restrict__ buf) {
    const int nj = planar ? 1 : L.Fy;
    const long long cnt = (long long)nj * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int j = (planar ? 0 : -1) + (int)(q / L.Fz), k = (int)(q % L.Fz) - 1;
        buf[q] = A[mg_at(L, i, j, k)];
    }
}
__global__ void __launch_bounds__(256) k_unpack_x(double *__restrict__ A, const LevelView L, const int i, const int planar, const double *__restrict__ buf) {
    const int nj = planar ? 1 : L.Fy;
    const long long cnt = (long long)nj * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int j = (planar ? 0 : -1) + (int)(q / L.Fz), k = (int)(q % L.Fz) - 1;
        A[mg_at(L, i, j, k)] = buf[q];
    }
}

/* Fused halo staging for the two-phase exchange (x faces first, then y rows that carry the fresh x pads):
 *   k_pack_x2          planes i = 1 and i = nx-2            -> send buffers of the left / right neighbour
 *   k_unpack_x_pack_y  receive buffers -> pads i = -1 / nx, and rows j = 1 / ny-2 (their i = -1 / nx entries taken from the
 *                      receive buffers, so no thread depends on another thread's pad write) -> y send buffers
 *   k_unpack_y2        receive buffers -> pads j = -1 / ny
 * Null pointers switch a side off. */
__global__ void __launch_bounds__(256) k_pack_x2(const double *__restrict__ A, const LevelView L, const int planar,
                                                 double *__restrict__ lo, double *__restrict__ hi) {
    const int nj = planar ? 1 : L.Fy;
    const long long cnt = (long long)nj * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int j = (planar ? 0 : -1) + (int)(q / L.Fz), k = (int)(q % L.Fz) - 1;
        if (lo) lo[q] = A[mg_at(L, 1, j, k)];
        if (hi) hi[q] = A[mg_at(L, L.nx - 2, j, k)];
    }
}
__global__ void __launch_bounds__(256) k_unpack_x_pack_y(double *__restrict__ A, const LevelView L, const int planar,
                                                         const double *__restrict__ xr_lo, const double *__restrict__ xr_hi,
                                                         double *__restrict__ ys_lo, double *__restrict__ ys_hi) {
    const int nj = planar ? 1 : L.Fy;
    const long long cx = (xr_lo || xr_hi) ? (long long)nj * L.Fz : 0;
    const long long cy = (ys_lo || ys_hi) ? (long long)L.Fx * L.Fz : 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cx + cy; q += (long long)gridDim.x * blockDim.x) {
        if (q < cx) {
            const int j = (planar ? 0 : -1) + (int)(q / L.Fz), k = (int)(q % L.Fz) - 1;
            if (xr_lo) A[mg_at(L, -1, j, k)] = xr_lo[q];
            if (xr_hi) A[mg_at(L, L.nx, j, k)] = xr_hi[q];
        } else {
            const long long w = q - cx;
            const int i = (int)(w / L.Fz) - 1, k = (int)(w % L.Fz) - 1;
            for (int side = 0; side < 2; side++) {
                double *dst = side ? ys_hi : ys_lo;
                if (!dst) continue;
                const int j = side ? L.ny - 2 : 1;
                double v;
                if (i == -1 && xr_lo) v = xr_lo[(long long)(j + 1) * L.Fz + (k + 1)];
                else if (i == L.nx && xr_hi) v = xr_hi[(long long)(j + 1) * L.Fz + (k + 1)];
                else v = A[mg_at(L, i, j, k)];
                dst[w] = v;
            }
        }
    }
}
__global__ void __launch_bounds__(256) k_unpack_y2(double *__restrict__ A, const LevelView L, const double *__restrict__ yr_lo,
                                                   const double *__restrict__ yr_hi) {
    const long long cnt = (long long)L.Fx * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / L.Fz) - 1, k = (int)(q % L.Fz) - 1;
        if (yr_lo) A[mg_at(L, i, -1, k)] = yr_lo[q];
        if (yr_hi) A[mg_at(L, i, L.ny, k)] = yr_hi[q];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Peer-memory halo exchange over NVLink (one process per GPU, buffers shared with CUDA IPC).  The three  */
/* kernels replace pack -> ncclSend/ncclRecv -> unpack of each phase with direct stores into the          */
/* neighbour's receive buffers:                                                                           */
/*   k_ipc_send_x   planes i = 1 / nx-2 -> left / right neighbour's x receive buffers, then their flags     */
/*   k_ipc_x_to_y   wait for my x flags; receive buffers -> pads i = -1 / nx; rows j = 1 / ny-2 (with the    */
/*                  fresh x pads taken from the receive buffers) -> front / back neighbour's y buffers, flags */
/*   k_ipc_recv_y   wait for my y flags; receive buffers -> pads j = -1 / ny                                */
/* Flags carry the exchange number (monotonic), receive buffers are double-buffered by its parity: a       */
/* neighbour can only be one exchange ahead (it needs my flag of exchange e+1 to finish e+1), so it never    */
/* overwrites a slot I have not consumed.  Data stores are made visible with __threadfence_system() before   */
/* the last CTA of the grid raises the remote flags; consumers read the buffers with L1-bypassing loads.     */
/* ------------------------------------------------------------------------------------------------ */
/* every CTA calls this after its stores; the last one to arrive raises the flags */
__device__ __forceinline__ void ipc_last_cta_signals(unsigned *counter, unsigned *flag_a, unsigned *flag_b, const unsigned epoch) {
#ifdef MG_FENCE_ALL
    __threadfence_system();
#endif
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();         /* one fence per CTA: cumulative over the stores the barrier ordered before it */
        const unsigned t = atomicAdd(counter, 1u);
        if (t == gridDim.x - 1) { *counter = 0u; ipc_signal(flag_a, epoch); ipc_signal(flag_b, epoch); }    /* st.release.sys */
    }
}
__device__ __forceinline__ double ld_cg_f64(const double *p) { return __ldcg(p); }

/* Addresses of one exchange.  Pointers are those of parity slot 0; slot 1 lies `slot_stride` doubles (buffers) / one
 * entry (flags) further.  The exchange number is *base + add: `base` is a device word the host (re)sets before a CUDA graph
 * of a whole V-cycle is replayed, `add` the position of the exchange inside the graph (or since the last reset), so the
 * same captured kernels carry fresh, dense exchange numbers on every replay. */
struct IpcPlan {
    double *send[4];            /* neighbour's receive buffer for my left / right x plane and front / back y row (nullptr: none) */
    unsigned *sflag[4];         /* the flag of that buffer, in the neighbour's arena */
    const double *recv[4];      /* my receive buffers: x lo, x hi, y lo, y hi */
    const unsigned *rflag[4];   /* my flags */
    long long slot_stride;
    const unsigned *base;
    unsigned add;
    unsigned *counter;          /* two arrival counters (x phase, y phase) of the multi-CTA kernels */
    unsigned *err;              /* device word: set when a wait timed out */
};
__device__ __forceinline__ unsigned ipc_epoch(const IpcPlan &P) { return *P.base + P.add; }
template <typename T>
__device__ __forceinline__ T *ipc_slot(T *p, const unsigned ep, const long long stride) { return p ? p + (ep & 1u) * stride : p; }

__global__ void __launch_bounds__(256) k_ipc_send_x(const double *__restrict__ A, const LevelView L, const int planar, const IpcPlan P) {
    const unsigned ep = ipc_epoch(P);
    double *dst_left = ipc_slot(P.send[0], ep, P.slot_stride), *dst_right = ipc_slot(P.send[1], ep, P.slot_stride);
    const int nj = planar ? 1 : L.Fy;
    const long long cnt = (long long)nj * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int j = (planar ? 0 : -1) + (int)(q / L.Fz), k = (int)(q % L.Fz) - 1;
        if (dst_left) dst_left[q] = A[mg_at(L, 1, j, k)];
        if (dst_right) dst_right[q] = A[mg_at(L, L.nx - 2, j, k)];
    }
    ipc_last_cta_signals(P.counter, ipc_slot(P.sflag[0], ep, 1), ipc_slot(P.sflag[1], ep, 1), ep);
}

__global__ void __launch_bounds__(256) k_ipc_x_to_y(double *__restrict__ A, const LevelView L, const int planar, const IpcPlan P) {
    const unsigned ep = ipc_epoch(P);
    if (threadIdx.x == 0) { ipc_wait_flag(ipc_slot(P.rflag[0], ep, 1), ep, P.err); ipc_wait_flag(ipc_slot(P.rflag[1], ep, 1), ep, P.err); }
    __syncthreads();
    const double *xr_lo = ipc_slot(P.recv[0], ep, P.slot_stride), *xr_hi = ipc_slot(P.recv[1], ep, P.slot_stride);
    double *dst_front = ipc_slot(P.send[2], ep, P.slot_stride), *dst_back = ipc_slot(P.send[3], ep, P.slot_stride);
    const int nj = planar ? 1 : L.Fy;
    const long long cx = (xr_lo || xr_hi) ? (long long)nj * L.Fz : 0;
    const long long cy = (dst_front || dst_back) ? (long long)L.Fx * L.Fz : 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cx + cy; q += (long long)gridDim.x * blockDim.x) {
        if (q < cx) {
            const int j = (planar ? 0 : -1) + (int)(q / L.Fz), k = (int)(q % L.Fz) - 1;
            if (xr_lo) A[mg_at(L, -1, j, k)] = ld_cg_f64(xr_lo + q);
            if (xr_hi) A[mg_at(L, L.nx, j, k)] = ld_cg_f64(xr_hi + q);
        } else {
            const long long w = q - cx;
            const int i = (int)(w / L.Fz) - 1, k = (int)(w % L.Fz) - 1;
            for (int side = 0; side < 2; side++) {
                double *dst = side ? dst_back : dst_front;
                if (!dst) continue;
                const int j = side ? L.ny - 2 : 1;
                double v;
                if (i == -1 && xr_lo) v = ld_cg_f64(xr_lo + (long long)(j + 1) * L.Fz + (k + 1));
                else if (i == L.nx && xr_hi) v = ld_cg_f64(xr_hi + (long long)(j + 1) * L.Fz + (k + 1));
                else v = A[mg_at(L, i, j, k)];
                dst[w] = v;
            }
        }
    }
    if (dst_front || dst_back) ipc_last_cta_signals(P.counter + 1, ipc_slot(P.sflag[2], ep, 1), ipc_slot(P.sflag[3], ep, 1), ep);
}

__global__ void __launch_bounds__(256) k_ipc_recv_y(double *__restrict__ A, const LevelView L, const IpcPlan P) {
    const unsigned ep = ipc_epoch(P);
    if (threadIdx.x == 0) { ipc_wait_flag(ipc_slot(P.rflag[2], ep, 1), ep, P.err); ipc_wait_flag(ipc_slot(P.rflag[3], ep, 1), ep, P.err); }
    __syncthreads();
    const double *yr_lo = ipc_slot(P.recv[2], ep, P.slot_stride), *yr_hi = ipc_slot(P.recv[3], ep, P.slot_stride);
    const long long cnt = (long long)L.Fx * L.Fz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / L.Fz) - 1, k = (int)(q % L.Fz) - 1;
        if (yr_lo) A[mg_at(L, i, -1, k)] = ld_cg_f64(yr_lo + q);
        if (yr_hi) A[mg_at(L, i, L.ny, k)] = ld_cg_f64(yr_hi + q);
    }
}

/* Small levels: the whole updatePads (+ the wall BCs of imposeBC) in ONE single-CTA launch.  On the latency-bound coarse
 * levels an exchange is a few thousand doubles, so the cost is launches and NVLink round trips, not bandwidth: the CTA
 * sends its x faces, raises the flags, waits for its own, unpacks, sends the y rows (with the fresh x pads), waits,
 * unpacks, and finishes with the face operations -- the same phases as the three kernels above, separated by
 * __syncthreads instead of launches. */
template <int DIM>
__global__ void __launch_bounds__(1024) k_ipc_exchange_cta(double *A, const LevelView L, const IpcPlan P,
                                                          const FaceModes self_x, const FaceModes self_y, const FaceModes bc) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const bool planar = (DIM == 2);
    const unsigned epoch = ipc_epoch(P);
    double *send[4];
    const double *recv[4];
    unsigned *sflag[4];
    const unsigned *rflag[4];
#pragma unroll
    for (int f = 0; f < 4; f++) {
        send[f] = ipc_slot(P.send[f], epoch, P.slot_stride); recv[f] = ipc_slot(P.recv[f], epoch, P.slot_stride);
        sflag[f] = ipc_slot(P.sflag[f], epoch, 1); rflag[f] = ipc_slot(P.rflag[f], epoch, 1);
    }
    const int nj = planar ? 1 : L.Fy;
    const int cx = nj * L.Fz, cy = planar ? 0 : L.Fx * L.Fz;
    faces_apply<DIM>(A, L, self_x, tid, nthr);                     /* periodic self-neighbour in x (no-op when all modes are NONE) */
    if (send[0] || send[1]) {
        for (int q = tid; q < cx; q += nthr) {
            const int j = (planar ? 0 : -1) + q / L.Fz, k = q % L.Fz - 1;
            if (send[0]) send[0][q] = A[mg_at(L, 1, j, k)];
            if (send[1]) send[1][q] = A[mg_at(L, L.nx - 2, j, k)];
        }
#ifdef MG_FENCE_ALL
        __threadfence_system();
#endif
        __syncthreads();
        if (tid == 0) { __threadfence_system(); ipc_signal(sflag[0], epoch); ipc_signal(sflag[1], epoch); }
    }
    if (tid == 0) { ipc_wait_flag(rflag[0], epoch, P.err); ipc_wait_flag(rflag[1], epoch, P.err); }
    __syncthreads();
    if (recv[0] || recv[1])
        for (int q = tid; q < cx; q += nthr) {
            const int j = (planar ? 0 : -1) + q / L.Fz, k = q % L.Fz - 1;
            if (recv[0]) A[mg_at(L, -1, j, k)] = ld_cg_f64(recv[0] + q);
            if (recv[1]) A[mg_at(L, L.nx, j, k)] = ld_cg_f64(recv[1] + q);
        }
    __syncthreads();
    if (DIM == 3) {
        faces_apply<DIM>(A, L, self_y, tid, nthr);
        if (send[2] || send[3]) {
            for (int q = tid; q < cy; q += nthr) {
                const int i = q / L.Fz - 1, k = q % L.Fz - 1;
                if (send[2]) send[2][q] = A[mg_at(L, i, 1, k)];
                if (send[3]) send[3][q] = A[mg_at(L, i, L.ny - 2, k)];
            }
#ifdef MG_FENCE_ALL
            __threadfence_system();
#endif
            __syncthreads();
            if (tid == 0) { __threadfence_system(); ipc_signal(sflag[2], epoch); ipc_signal(sflag[3], epoch); }
        }
        if (tid == 0) { ipc_wait_flag(rflag[2], epoch, P.err); ipc_wait_flag(rflag[3], epoch, P.err); }
        __syncthreads();
        if (recv[2] || recv[3])
            for (int q = tid; q < cy; q += nthr) {
                const int i = q / L.Fz - 1, k = q % L.Fz - 1;
                if (recv[2]) A[mg_at(L, i, -1, k)] = ld_cg_f64(recv[2] + q);
                if (recv[3]) A[mg_at(L, i, L.ny, k)] = ld_cg_f64(recv[3] + q);
            }
        __syncthreads();
    }
    faces_apply<DIM>(A, L, bc, tid, nthr);                          /* wall BCs of imposeBC (all NONE for a bare updatePads) */
}
__global__ void k_set_u32(unsigned *p, const unsigned v) { *p = v; }

/* ------------------------------------------------------------------------------------------------ */
/* full-weighting restriction (d3:270-297, d2:222-235): thread per coarse point, sums in source order */
/* ------------------------------------------------------------------------------------------------ */
template <int DIM>
__global__ void __launch_bounds__(256) k_coarsen(const double *__restrict__ T, const LevelView Lf, double *__restrict__ Rc, const LevelView Lc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = (DIM == 3) ? blockIdx.y * blockDim.y + threadIdx.y : 0;
    const int i = blockIdx.z;
    if (k >= Lc.nz || j >= Lc.ny || i >= Lc.nx) return;
    const int i2 = i * 2, j2 = j * 2, k2 = k * 2;
#define TF(a, b, c) T[mg_at(Lf, (a), (b), (c))]
    if (DIM == 3) {
        const double facePoints = (TF(i2 + 1, j2, k2) + TF(i2 - 1, j2, k2) +
                                   TF(i2, j2 + 1, k2) + TF(i2, j2 - 1, k2) +
                                   TF(i2, j2, k2 + 1) + TF(i2, j2, k2 - 1)) * 0.0625;
        const double edgePoints = (TF(i2 + 1, j2 + 1, k2) + TF(i2 + 1, j2 - 1, k2) +
                                   TF(i2 - 1, j2 - 1, k2) + TF(i2 - 1, j2 + 1, k2) +
                                   TF(i2, j2 + 1, k2 + 1) + TF(i2, j2 - 1, k2 + 1) +
                                   TF(i2, j2 - 1, k2 - 1) + TF(i2, j2 + 1, k2 - 1) +
                                   TF(i2 + 1, j2, k2 + 1) + TF(i2 + 1, j2, k2 - 1) +
                                   TF(i2 - 1, j2, k2 - 1) + TF(i2 - 1, j2, k2 + 1)) * 0.03125;
        const double vertPoints = (TF(i2 + 1, j2 + 1, k2 + 1) +
                                   TF(i2 + 1, j2 + 1, k2 - 1) +
                                   TF(i2 + 1, j2 - 1, k2 + 1) +
                                   TF(i2 - 1, j2 + 1, k2 + 1) +
                                   TF(i2 + 1, j2 - 1, k2 - 1) +
                                   TF(i2 - 1, j2 + 1, k2 - 1) +
                                   TF(i2 - 1, j2 - 1, k2 + 1) +
                                   TF(i2 - 1, j2 - 1, k2 - 1)) * 0.015625;
        Rc[mg_at(Lc, i, j, k)] = facePoints + edgePoints + vertPoints + TF(i2, j2, k2) * 0.125;
    } else {
        const double facePoints = (TF(i2 + 1, 0, k2) + TF(i2 - 1, 0, k2) +
                                   TF(i2, 0, k2 + 1) + TF(i2, 0, k2 - 1)) * 0.125;
        const double vertPoints = (TF(i2 + 1, 0, k2 + 1) +
                                   TF(i2 + 1, 0, k2 - 1) +
                                   TF(i2 - 1, 0, k2 + 1) +
                                   TF(i2 - 1, 0, k2 - 1)) * 0.0625;
        Rc[mg_at(Lc, i, 0, k)] = facePoints + vertPoints + TF(i2, 0, k2) * 0.25;
    }
#undef TF
}

/* The 3-D restriction as an x-marching kernel: a thread owns the coarse column (J, K) and walks I.  The nine fine values
 * of plane 2I+1 it loaded for coarse point I are the plane 2(I+1)-1 values of the next point, so every step loads two
 * new fine planes (3 rows each: one aligned 128-bit load for (k2, k2+1) plus one 64-bit load for k2-1) instead of 27
 * scalars, and all of a step's loads are issued before the sums.  Sums in the order of d3:270-297. */
__global__ void __launch_bounds__(256) k_coarsen_march(const double *__restrict__ T, const LevelView Lf, double *__restrict__ Rc,
                                                       const LevelView Lc, const int ichunk) {
    const int K = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (K >= Lc.nz || J >= Lc.ny) return;
    const int I0 = blockIdx.z * ichunk, I1 = min(I0 + ichunk, Lc.nx);
    const long long fplane = Lf.plane;
    const int fp = Lf.pitch;
    /* fine row j2 = 2J, column k2 = 2K of x-plane index 0 (i = -1) */
    const double *t = T + (long long)(2 * J + 1) * fp + (2 * K + 1 + MG_LEAD);
    double *rc = Rc + (long long)(J + 1) * Lc.pitch + (K + 1 + MG_LEAD);
    double A[3][3], B[3][3], C[3][3];
#define LOAD_PLANE(M, base)                                                          \
    {                                                                                \
        _Pragma("unroll") for (int r = 0; r < 3; r++) {                              \
            const double *row = (base) + (long long)(r - 1) * fp;                    \
            const double2 v = ld2(row);                                              \
            M[r][0] = row[-1]; M[r][1] = v.x; M[r][2] = v.y;                         \
        }                                                                            \
    }
    LOAD_PLANE(A, t + (long long)(2 * I0) * fplane);                /* fine plane 2*I0 - 1 */
    for (int I = I0; I < I1; ++I) {
        const double *b = t + (long long)(2 * I + 1) * fplane;      /* fine plane 2I */
        LOAD_PLANE(B, b);
        LOAD_PLANE(C, b + fplane);
        const double facePoints = (C[1][1] + A[1][1] + B[2][1] + B[0][1] + B[1][2] + B[1][0]) * 0.0625;
        const double edgePoints = (C[2][1] + C[0][1] + A[0][1] + A[2][1] + B[2][2] + B[0][2] + B[0][0] + B[2][0] +
                                   C[1][2] + C[1][0] + A[1][0] + A[1][2]) * 0.03125;
        const double vertPoints = (C[2][2] + C[2][0] + C[0][2] + A[2][2] + C[0][0] + A[2][0] + A[0][2] + A[0][0]) * 0.015625;
        rc[(long long)(I + 1) * Lc.plane] = facePoints + edgePoints + vertPoints + B[1][1] * 0.125;
#pragma unroll
        for (int r = 0; r < 3; r++) { A[r][0] = C[r][0]; A[r][1] = C[r][1]; A[r][2] = C[r][2]; }
    }
#undef LOAD_PLANE
}

/* ------------------------------------------------------------------------------------------------ */
/* prolongation fused with "pressureData += smoothedPres" (d3:310-365 + p:271; d2:248-270).            */
/* The fine array still holds the pre-coarsening iterate (it IS smoothedPres), so                      */
/*    fine[idx] = interp + fine[idx]   ( = (0 + interp) + smoothedPres, addition commutes exactly ).   */
/* Algorithmic traffic 17 B per fine point.                                                           */
/* ------------------------------------------------------------------------------------------------ */
template <int DIM>
__global__ void __launch_bounds__(256) k_prolong_add(const double *__restrict__ C, const LevelView Lc, double *__restrict__ Pf, const LevelView Lf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = (DIM == 3) ? blockIdx.y * blockDim.y + threadIdx.y : 0;
    const int i = blockIdx.z;
    if (k >= Lf.nz || j >= Lf.ny || i >= Lf.nx) return;
    const int i2 = i / 2, j2 = j / 2, k2 = k / 2;
    const bool oi = i & 1, oj = (DIM == 3) ? (j & 1) : false, ok = k & 1;
#define PC(a, b, c) C[mg_at(Lc, (a), (b), (c))]
    double v;
    if (DIM == 3) {
        if (oi) {
            if (oj) {
                if (ok) v = (PC(i2, j2, k2) +
                             PC(i2 + 1, j2, k2) + PC(i2, j2 + 1, k2) + PC(i2, j2, k2 + 1) +
                             PC(i2 + 1, j2 + 1, k2) + PC(i2 + 1, j2, k2 + 1) + PC(i2, j2 + 1, k2 + 1) +
                             PC(i2 + 1, j2 + 1, k2 + 1)) / 8.0;
                else    v = (PC(i2, j2, k2) + PC(i2, j2 + 1, k2) + PC(i2 + 1, j2, k2) + PC(i2 + 1, j2 + 1, k2)) / 4.0;
            } else {
                if (ok) v = (PC(i2, j2, k2) + PC(i2, j2, k2 + 1) + PC(i2 + 1, j2, k2) + PC(i2 + 1, j2, k2 + 1)) / 4.0;
                else    v = (PC(i2, j2, k2) + PC(i2 + 1, j2, k2)) / 2.0;
            }
        } else {
            if (oj) {
                if (ok) v = (PC(i2, j2, k2) + PC(i2, j2, k2 + 1) + PC(i2, j2 + 1, k2) + PC(i2, j2 + 1, k2 + 1)) / 4.0;
                else    v = (PC(i2, j2, k2) + PC(i2, j2 + 1, k2)) / 2.0;
            } else {
                if (ok) v = (PC(i2, j2, k2) + PC(i2, j2, k2 + 1)) / 2.0;
                else    v = PC(i2, j2, k2);
            }
        }
    } else {
        if (oi) {
            if (ok) v = (PC(i2, 0, k2) + PC(i2, 0, k2 + 1) + PC(i2 + 1, 0, k2) + PC(i2 + 1, 0, k2 + 1)) / 4.0;
            else    v = (PC(i2, 0, k2) + PC(i2 + 1, 0, k2)) / 2.0;
        } else {
            if (ok) v = (PC(i2, 0, k2) + PC(i2, 0, k2 + 1)) / 2.0;
            else    v = PC(i2, 0, k2);
        }
    }
#undef PC
    const long long idx = mg_at(Lf, i, j, k);
    Pf[idx] = v + Pf[idx];
}

/* The same operator as an x-marching kernel (used for every level with >= 2 coarse points along z per warp lane).
 * A thread owns the coarse cell (j2, k2) and walks i2: the four coarse values of plane i2 it needs stay in registers
 * and become the "minus" plane of the next step, every fine row is touched with one aligned 128-bit load and one
 * 128-bit store per thread (fine k = 2*k2 sits on a 16-byte boundary), and a warp covers 512 contiguous bytes of each
 * of the four fine rows (2i2 | 2i2+1, 2j2 | 2j2+1).  Sums are left-associated in the order of d3:310-365 / d2:248-270. */
template <int DIM>
__global__ void __launch_bounds__(256) k_prolong_march(const double *__restrict__ C, const LevelView Lc, double *Pf,
                                                       const LevelView Lf, const int ichunk) {
    const int k2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int j2 = (DIM == 3) ? (blockIdx.y * blockDim.y + threadIdx.y) : 0;
    if (k2 >= Lc.nz || j2 >= Lc.ny) return;
    const int i0 = blockIdx.z * ichunk;                 /* coarse planes [i0, i1) */
    const int i1 = min(i0 + ichunk, Lc.nx);
    const bool hk = (k2 + 1 < Lc.nz);                   /* fine k = 2*k2+1 exists */
    const bool hj = (DIM == 3) && (j2 + 1 < Lc.ny);     /* fine j = 2*j2+1 exists */
    const double *c = C + (long long)(j2 + 1) * Lc.pitch + (k2 + 1 + MG_LEAD);
    double *f = Pf + (long long)((DIM == 3 ? 2 * j2 : 0) + 1) * Lf.pitch + (2 * k2 + 1 + MG_LEAD);
    const int cp = Lc.pitch, fp = Lf.pitch;
    const long long fplane = Lf.plane;
    /* coarse values: [di][dj][dk] relative to (i2, j2, k2); entries beyond the array edge are never used */
    double a00, a01 = 0.0, a10 = 0.0, a11 = 0.0;        /* plane i2 */
    {
        const double *q = c + (long long)(i0 + 1) * Lc.plane;
        a00 = q[0];
        if (hk) a01 = q[1];
        if (hj) { a10 = q[cp]; if (hk) a11 = q[cp + 1]; }
    }
    /* The fine array is updated in place, so the compiler may not move a step's loads above the previous step's stores;
     * the loads of the NEXT step are therefore issued by hand before the stores of this one (two steps of traffic in flight). */
    auto load_rows = [&](const int i2, double2 (&v)[4]) {           /* rows (2i2, 2j2), (2i2, 2j2+1), (2i2+1, 2j2), (2i2+1, 2j2+1) */
        const double *fe = f + (long long)(2 * i2 + 1) * fplane;
        const bool hi = (i2 + 1 < Lc.nx);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            v[r] = make_double2(0.0, 0.0);
            const bool ok = ((r & 1) == 0 || hj) && (r < 2 || hi);
            if (ok) {
                const double *g = fe + (r >> 1) * fplane + (r & 1) * fp;
                if (hk) v[r] = ld2(g); else v[r].x = g[0];
            }
        }
    };
    double2 cur[4], nxt[4];
    load_rows(i0, cur);
    for (int i2 = i0; i2 < i1; ++i2) {
        const bool hi = (i2 + 1 < Lc.nx);               /* fine i = 2*i2+1 exists */
        double b00 = 0.0, b01 = 0.0, b10 = 0.0, b11 = 0.0;   /* plane i2+1 */
        if (hi) {
            const double *q = c + (long long)(i2 + 2) * Lc.plane;
            b00 = q[0];
            if (hk) b01 = q[1];
            if (hj) { b10 = q[cp]; if (hk) b11 = q[cp + 1]; }
        }
        if (i2 + 1 < i1) load_rows(i2 + 1, nxt);
        double *fe = f + (long long)(2 * i2 + 1) * fplane;           /* fine plane 2*i2 */
        double *fo = fe + fplane;                                     /* fine plane 2*i2+1 */
        double2 v;
        v.x = a00;
        v.y = (a00 + a01) / 2.0;
        if (hk) st2(fe, make_double2(v.x + cur[0].x, v.y + cur[0].y)); else fe[0] = v.x + cur[0].x;
        if (hj) {
            v.x = (a00 + a10) / 2.0;
            v.y = (a00 + a01 + a10 + a11) / 4.0;
            if (hk) st2(fe + fp, make_double2(v.x + cur[1].x, v.y + cur[1].y)); else fe[fp] = v.x + cur[1].x;
        }
        if (hi) {
            v.x = (a00 + b00) / 2.0;
            v.y = (a00 + a01 + b00 + b01) / 4.0;
            if (hk) st2(fo, make_double2(v.x + cur[2].x, v.y + cur[2].y)); else fo[0] = v.x + cur[2].x;
            if (hj) {
                v.x = (a00 + a10 + b00 + b10) / 4.0;
                v.y = (a00 + b00 + a10 + a01 + b10 + b01 + a11 + b11) / 8.0;
                if (hk) st2(fo + fp, make_double2(v.x + cur[3].x, v.y + cur[3].y)); else fo[fp] = v.x + cur[3].x;
            }
        }
        a00 = b00; a01 = b01; a10 = b10; a11 = b11;
#pragma unroll
        for (int r = 0; r < 4; r++) cur[r] = nxt[r];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* exact lexicographic Gauss-Seidel as a hyperplane wavefront (d3:137-150, d2:119-128).                */
/* Point (i,j,k) reads the NEW values of (i-1,j,k), (i,j-1,k), (i,j,k-1) and the OLD values of the +1   */
/* neighbours; all points with i+j+k = h are mutually independent and see exactly the operands the      */
/* sequential i->j->k loop would give them, so processing h = 0,1,2,... in order is bit-identical.      */
/* ------------------------------------------------------------------------------------------------ */
template <int DIM>
__device__ __forceinline__ void gs_point(double *P, const double *R, const LevelView &L, const int i, const int j, const int k) {
    const long long c = mg_at(L, i, j, k);
    if (DIM == 3) {
        const double cx2 = L.hyhz * L.xix2[i], cx1 = L.hyhz * L.xixx[i];
        const double cy2 = L.hzhx * L.ety2[j], cy1 = L.hzhx * L.etyy[j];
        const double cz2 = L.hxhy * L.ztz2[k], cz1 = L.hxhy * L.ztzz[k];
        P[c] = relax3(P[c + L.plane], P[c - L.plane], P[c + L.pitch], P[c - L.pitch], P[c + 1], P[c - 1], R[c],
                      cx2, cx1, L.hx, cy2, cy1, L.hy, cz2, cz1, L.hz, 2.0 * L.hxhyhz, 4.0 * (cx2 + cy2 + cz2));
    } else {
        const double cx2 = L.hz2 * L.xix2[i], cx1 = L.hz2 * L.xixx[i];
        const double cz2 = L.hx2 * L.ztz2[k], cz1 = L.hx2 * L.ztzz[k];
        P[c] = relax2(P[c + L.plane], P[c - L.plane], P[c + 1], P[c - 1], R[c],
                      cx2, cx1, L.hx, cz2, cz1, L.hz, 2.0 * L.hzhx, 4.0 * (cx2 + cz2));
    }
}

/* one hyperplane: threads over (j fast, i slow) for 3-D, over i for 2-D */
template <int DIM>
__global__ void __launch_bounds__(256) k_gs_plane(double *P, const double *__restrict__ R, const LevelView L, const int h, const int ilo) {
    if (DIM == 3) {
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        const int i = ilo + blockIdx.y * blockDim.y + threadIdx.y;
        const int k = h - i - j;
        if (j >= L.ny || i >= L.nx || k < 0 || k >= L.nz) return;
        gs_point<3>(P, R, L, i, j, k);
    } else {
        const int i = ilo + blockIdx.x * blockDim.x + threadIdx.x;
        const int k = h - i;
        if (i >= L.nx || k < 0 || k >= L.nz) return;
        gs_point<2>(P, R, L, i, 0, k);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* residual max for the coarsest-level solve (d3:213-231, d2:169-183): max(|rhs - L p|, -1e-10)        */
/* ------------------------------------------------------------------------------------------------ */
template <int DIM>
__device__ __forceinline__ double resid_abs_point(const double *P, const double *R, const LevelView &L, const int i, const int j, const int k) {
    const long long c = mg_at(L, i, j, k);
    if (DIM == 3) {
        return fabs(R[c] - lap3(P[c], P[c + L.plane], P[c - L.plane], P[c + L.pitch], P[c - L.pitch], P[c + 1], P[c - 1],
                                L.xix2[i], L.xixx[i], L.ety2[j], L.etyy[j], L.ztz2[k], L.ztzz[k],
                                L.hx * L.hx, 2.0 * L.hx, L.hy * L.hy, 2.0 * L.hy, L.hz * L.hz, 2.0 * L.hz));
    } else {
        return fabs(R[c] - lap2(P[c], P[c + L.plane], P[c - L.plane], P[c + 1], P[c - 1],
                                L.xix2[i], L.xixx[i], L.ztz2[k], L.ztzz[k],
                                L.hx * L.hx, 2.0 * L.hx, L.hz * L.hz, 2.0 * L.hz));
    }
}

template <int DIM>
__global__ void __launch_bounds__(256) k_resid_max(const double *__restrict__ P, const double *__restrict__ R, const LevelView L, double *__restrict__ part) {
    __shared__ double sm[32];
    const long long cnt = (long long)L.nx * L.ny * L.nz;
    double m = -1.0e-10;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(q % L.nz);
        const int j = (int)((q / L.nz) % L.ny);
        const int i = (int)(q / ((long long)L.nz * L.ny));
        m = fmax(m, resid_abs_point<DIM>(P, R, L, i, j, k));
    }
    m = block_reduce<true>(m, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = m;
}

/* ------------------------------------------------------------------------------------------------ */
/* Coarsest-level solve, single rank: the whole `while(true)` of d3:190-247 / d2:150-199 in ONE CTA.   */
/* imposeBC (self-neighbour pad copies x then y, then wall BCs x, y, z), exact GS wavefront, residual   */
/* max and the convergence test all stay on the device; the iteration count is written to *iters_out.  */
/* ------------------------------------------------------------------------------------------------ */
struct BcPlan {
    int pad_x_lo, pad_x_hi, pad_y_lo, pad_y_hi;   /* FACE_PERIODIC where the neighbour is this rank itself, else FACE_NONE */
    int bc_x_lo, bc_x_hi, bc_y_lo, bc_y_hi, bc_z_lo, bc_z_hi;
    const double *xwall, *ywall, *zwall;
    int planar;
};

template <int DIM>
__device__ void cta_impose_bc(double *P, const LevelView &L, const BcPlan &B) {
    const long long tid = threadIdx.x, nthr = blockDim.x;
    if (B.pad_x_lo | B.pad_x_hi) { face_op<0>(P, L, B.pad_x_lo, B.pad_x_hi, nullptr, B.planar, tid, nthr); __syncthreads(); }
    if (DIM == 3 && (B.pad_y_lo | B.pad_y_hi)) { face_op<1>(P, L, B.pad_y_lo, B.pad_y_hi, nullptr, B.planar, tid, nthr); __syncthreads(); }
    if (B.bc_x_lo | B.bc_x_hi) { face_op<0>(P, L, B.bc_x_lo, B.bc_x_hi, B.xwall, B.planar, tid, nthr); __syncthreads(); }
    if (DIM == 3 && (B.bc_y_lo | B.bc_y_hi)) { face_op<1>(P, L, B.bc_y_lo, B.bc_y_hi, B.ywall, B.planar, tid, nthr); __syncthreads(); }
    if (B.bc_z_lo | B.bc_z_hi) { face_op<2>(P, L, B.bc_z_lo, B.bc_z_hi, B.zwall, B.planar, tid, nthr); __syncthreads(); }
}

template <int DIM>
__global__ void __launch_bounds__(1024) k_coarse_solve(double *P, const double *R, const LevelView L, const BcPlan B,
                                                      const double tol, const int maxCount, int *iters_out) {
    __shared__ double sm[32];
    __shared__ double s_max;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int H = L.nx + L.ny + L.nz - 2;
    const int nij = L.nx * L.ny;
    const long long cnt = (long long)nij * L.nz;
    int iter = 0;
    while (true) {
        cta_impose_bc<DIM>(P, L, B);
        for (int h = 0; h < H; ++h) {
            for (int q = tid; q < nij; q += nthr) {
                const int i = q / L.ny, j = q % L.ny;
                const int k = h - i - j;
                if (k >= 0 && k < L.nz) gs_point<DIM>(P, R, L, i, j, k);
            }
            __syncthreads();
        }
        double m = -1.0e-10;
        for (long long q = tid; q < cnt; q += nthr) {
            const int k = (int)(q % L.nz);
            const int j = (int)((q / L.nz) % L.ny);
            const int i = (int)(q / ((long long)L.nz * L.ny));
            m = fmax(m, resid_abs_point<DIM>(P, R, L, i, j, k));
        }
        m = block_reduce<true>(m, sm);
        if (tid == 0) s_max = m;
        __syncthreads();
        const double gmax = s_max;
        __syncthreads();
        if (gmax < tol) break;
        iter += 1;
        if (iter > maxCount) break;
    }
    cta_impose_bc<DIM>(P, L, B);
    if (tid == 0) *iters_out = iter;
}

/* ------------------------------------------------------------------------------------------------ */
/* Coarsest-level solve, several ranks: agglomerated.  The reference iterates {imposeBC (halo exchange   */
/* + wall BCs); per-rank lexicographic GS sweep; global MAX of the residual} until the MAX drops under    */
/* mgTolerance (d3:190-247) -- hundreds of latency-bound exchange + all-reduce rounds on a grid of a few  */
/* points per rank.  Here every rank all-gathers the (tiny) coarsest arrays of all ranks once, and ONE    */
/* CTA replays the whole loop on the nb rank-blocks exactly as the nb reference ranks would: each block   */
/* keeps its own pads, exchanges faces with its neighbour blocks (x phase, then y phase carrying the      */
/* fresh x pads), applies its own wall BCs, and is swept by its own block-GS; the convergence test uses   */
/* the MAX over all blocks.  Every rank computes the same bits and keeps its own block, so no scatter     */
/* and no per-iteration collective is needed.                                                            */
/* ------------------------------------------------------------------------------------------------ */
struct CoarseBlock {
    int nbr[4];                     /* face neighbours (block ids) left, right, front, back; -1 = none */
    int wall[4];                    /* 1 where that face is a physical non-periodic wall of the domain */
    const double *xixx, *xix2, *etyy, *ety2;    /* this block's x / y metric arrays of the coarsest level (index base -1 applied) */
};

template <int DIM>
__global__ void __launch_bounds__(1024) k_coarse_solve_blocks(double *P, const double *R, const long long stride, const int nblk,
                                                             const CoarseBlock *__restrict__ tab, const LevelView L,
                                                             const int wall_mode, const int z_lo_mode, const int z_hi_mode,
                                                             const double tol, const int maxCount, int *iters_out) {
    __shared__ double sm[32];
    __shared__ double s_max;
    __shared__ CoarseBlock s_tab[16];
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int q = tid; q < nblk; q += nthr) s_tab[q] = tab[q];
    __syncthreads();
    const bool planar = (DIM == 2);
    const int H = L.nx + L.ny + L.nz - 2;
    const int nij = L.nx * L.ny;
    const long long cnt = (long long)nij * L.nz;
    const int jlo = planar ? 0 : -1, nj = planar ? 1 : L.Fy;
    const int xE = L.nx - 1, yE = L.ny - 1;
    int iter = 0;
    while (true) {
        /* ---- imposeBC of every block: updatePads x phase, y phase, then walls x, y, z (d3:575-668) ---- */
        {
            const long long per = (long long)nj * L.Fz;
            for (long long q = tid; q < per * nblk; q += nthr) {
                const int b = (int)(q / per);
                const long long w = q % per;
                const int j = jlo + (int)(w / L.Fz), k = (int)(w % L.Fz) - 1;
                double *A = P + b * stride;
                const int nl = s_tab[b].nbr[0], nr = s_tab[b].nbr[1];
                if (nl >= 0) A[mg_at(L, -1, j, k)] = P[nl * stride + mg_at(L, xE - 1, j, k)];
                if (nr >= 0) A[mg_at(L, xE + 1, j, k)] = P[nr * stride + mg_at(L, 1, j, k)];
            }
            __syncthreads();
        }
        if (DIM == 3) {
            const long long per = (long long)L.Fx * L.Fz;
            for (long long q = tid; q < per * nblk; q += nthr) {
                const int b = (int)(q / per);
                const long long w = q % per;
                const int i = (int)(w / L.Fz) - 1, k = (int)(w % L.Fz) - 1;
                double *A = P + b * stride;
                const int nf = s_tab[b].nbr[2], nk = s_tab[b].nbr[3];
                if (nf >= 0) A[mg_at(L, i, -1, k)] = P[nf * stride + mg_at(L, i, yE - 1, k)];
                if (nk >= 0) A[mg_at(L, i, yE + 1, k)] = P[nk * stride + mg_at(L, i, 1, k)];
            }
            __syncthreads();
        }
        for (int b = 0; b < nblk; ++b) {
            const int lo = s_tab[b].wall[0] ? wall_mode : FACE_NONE, hi = s_tab[b].wall[1] ? wall_mode : FACE_NONE;
            if (lo | hi) face_op<0>(P + b * stride, L, lo, hi, nullptr, planar, tid, nthr);
        }
        __syncthreads();
        if (DIM == 3) {
            for (int b = 0; b < nblk; ++b) {
                const int lo = s_tab[b].wall[2] ? wall_mode : FACE_NONE, hi = s_tab[b].wall[3] ? wall_mode : FACE_NONE;
                if (lo | hi) face_op<1>(P + b * stride, L, lo, hi, nullptr, planar, tid, nthr);
            }
            __syncthreads();
        }
        for (int b = 0; b < nblk; ++b) face_op<2>(P + b * stride, L, z_lo_mode, z_hi_mode, nullptr, planar, tid, nthr);
        __syncthreads();
        if (iter < 0) break;                            /* trailing imposeBC done */

        /* ---- block Gauss-Seidel: hyperplane wavefront inside every block, all blocks at once ---- */
        for (int h = 0; h < H; ++h) {
            for (int q = tid; q < nij * nblk; q += nthr) {
                const int b = q / nij, w = q % nij;
                const int i = w / L.ny, j = w % L.ny;
                const int k = h - i - j;
                if (k >= 0 && k < L.nz) {
                    LevelView Lb = L;
                    Lb.xixx = s_tab[b].xixx; Lb.xix2 = s_tab[b].xix2; Lb.etyy = s_tab[b].etyy; Lb.ety2 = s_tab[b].ety2;
                    gs_point<DIM>(P + b * stride, R + b * stride, Lb, i, j, k);
                }
            }
            __syncthreads();
        }
        double m = -1.0e-10;
        for (long long q = tid; q < cnt * nblk; q += nthr) {
            const int b = (int)(q / cnt);
            const long long w = q % cnt;
            const int k = (int)(w % L.nz);
            const int j = (int)((w / L.nz) % L.ny);
            const int i = (int)(w / ((long long)L.nz * L.ny));
            LevelView Lb = L;
            Lb.xixx = s_tab[b].xixx; Lb.xix2 = s_tab[b].xix2; Lb.etyy = s_tab[b].etyy; Lb.ety2 = s_tab[b].ety2;
            m = fmax(m, resid_abs_point<DIM>(P + b * stride, R + b * stride, Lb, i, j, k));
        }
        m = block_reduce<true>(m, sm);
        if (tid == 0) s_max = m;
        __syncthreads();
        const double gmax = s_max;
        __syncthreads();
        bool stop = gmax < tol;
        if (!stop) { iter += 1; stop = iter > maxCount; }
        if (stop) { if (tid == 0) *iters_out = iter; iter = -1; }      /* one more pass through imposeBC, then out */
    }
}

/* The same loop with the blocks resident in SHARED memory (used whenever they fit; nblk = 1 is the single-rank solve).
 * Per block: pressure and rhs in compact padded layout, the stencil diagonal and its correctly rounded reciprocal per core
 * point (the GS update divides by the diagonal: div_inv reproduces the IEEE quotient from the stored reciprocal), and the
 * coefficient products of every axis.  All index arithmetic is done ONCE: the five phases of imposeBC (x pads, y pads,
 * x walls, y walls, z) become (destination, source) lists, the GS wavefront and the residual pass walk a column table.
 * An iteration is then ~H short shared-memory phases (H = nx+ny+nz-2 hyperplanes) instead of chains of global loads,
 * integer divisions and an IEEE division per point. */
struct CoarseCol { int off, s, pt, xi, yj, pad; };   /* sP offset of (b,i,j,k=0); i+j; point index of k=0; b*nx+i; b*ny+j (24 bytes) */

template <int DIM>
__global__ void __launch_bounds__(256) k_coarse_solve_smem(double *Pg, const double *Rg, const long long stride, const int nblk,
                                                          const CoarseBlock *__restrict__ tab, const LevelView L,
                                                          const int wall_mode, const int z_lo_mode, const int z_hi_mode, const int p2,
                                                          const double tol, const int maxCount, int *iters_out) {
    extern __shared__ double cs_sm[];
    __shared__ double sm[32];
    __shared__ double s_max;
    __shared__ CoarseBlock s_tab[16];
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int q = tid; q < nblk; q += nthr) s_tab[q] = tab[q];
    const int Fx = L.Fx, Fy = (DIM == 2) ? 1 : L.Fy, Fz = L.Fz;      /* 2-D: only the live j = 0 row is kept */
    const int jo = (DIM == 2) ? 0 : 1;                               /* row offset of j = 0 */
    const int vol = Fx * Fy * Fz;
    const int nx = L.nx, ny = L.ny, nz = L.nz, nij = nx * ny, cnt = nij * nz, ncol = nblk * nij;
    double *sP = cs_sm, *sR = sP + nblk * vol, *sD = sR + nblk * vol, *sI = sD + nblk * cnt;
    double *cx2 = sI + nblk * cnt, *cx1 = cx2 + nblk * nx, *cy2 = cx1 + nblk * nx, *cy1 = cy2 + nblk * ny;
    double *cz2 = cy1 + nblk * ny, *cz1 = cz2 + nz, *mx2 = cz1 + nz, *mx1 = mx2 + nblk * nx, *my2 = mx1 + nblk * nx, *my1 = my2 + nblk * ny;
    double *mz2 = my1 + nblk * ny, *mz1 = mz2 + nz;
    /* integer tables behind the doubles */
    const int nX = 2 * nblk * Fy * Fz, nY = (DIM == 3) ? 2 * nblk * Fx * Fz : 0, nZ = 2 * nblk * Fx * Fy;
    CoarseCol *cols = reinterpret_cast<CoarseCol *>(mz1 + nz);
    int2 *lx = reinterpret_cast<int2 *>(cols + ncol), *ly = lx + nX, *lwx = ly + nY, *lwy = lwx + nX, *lz = lwy + nY;
#define CS_AT(i, j, k) ((((i) + 1) * Fy + ((j) + jo)) * Fz + ((k) + 1))
    __syncthreads();
    const int xE = nx - 1, yE = ny - 1, zE = nz - 1;
    const int sj = Fz, si = Fy * Fz;
    for (int q = tid; q < nblk * vol; q += nthr) {
        const int b = q / vol, w = q % vol;
        const int k = w % Fz - 1, j = (w / Fz) % Fy - jo, i = w / (Fz * Fy) - 1;
        sP[q] = Pg[b * stride + mg_at(L, i, j, k)];
        sR[q] = Rg[b * stride + mg_at(L, i, j, k)];
    }
    const double kxc = (DIM == 3) ? L.hyhz : L.hz2, kyc = L.hzhx, kzc = (DIM == 3) ? L.hxhy : L.hx2;
    for (int q = tid; q < nblk * nx; q += nthr) { const int b = q / nx, i = q % nx; mx2[q] = s_tab[b].xix2[i]; mx1[q] = s_tab[b].xixx[i]; cx2[q] = kxc * mx2[q]; cx1[q] = kxc * mx1[q]; }
    if (DIM == 3)
        for (int q = tid; q < nblk * ny; q += nthr) { const int b = q / ny, j = q % ny; my2[q] = s_tab[b].ety2[j]; my1[q] = s_tab[b].etyy[j]; cy2[q] = kyc * my2[q]; cy1[q] = kyc * my1[q]; }
    for (int q = tid; q < nz; q += nthr) { mz2[q] = L.ztz2[q]; mz1[q] = L.ztzz[q]; cz2[q] = kzc * mz2[q]; cz1[q] = kzc * mz1[q]; }
    for (int q = tid; q < ncol; q += nthr) {
        const int b = q / nij, w = q % nij, i = w / ny, j = w % ny;
        CoarseCol c;
        c.off = b * vol + CS_AT(i, j, 0); c.s = i + j; c.pt = b * cnt + w * nz; c.xi = b * nx + i; c.yj = b * ny + j; c.pad = 0;
        cols[q] = c;
    }
    /* imposeBC lists: entry = (destination, source) offsets into sP; source < 0 encodes "negate" as ~source; destination -1 = none */
    const bool ng = (wall_mode == FACE_NEG_MIRROR);
    for (int q = tid; q < nblk * Fy * Fz; q += nthr) {
        const int b = q / (Fy * Fz), w = q % (Fy * Fz), o = b * vol + w;
        const int nl = s_tab[b].nbr[0], nr = s_tab[b].nbr[1];
        lx[2 * q] = (nl >= 0) ? make_int2(o, nl * vol + xE * si + w) : make_int2(-1, 0);                 /* i = -1 <- neighbour's i = xE-1 */
        lx[2 * q + 1] = (nr >= 0) ? make_int2(o + (xE + 2) * si, nr * vol + 2 * si + w) : make_int2(-1, 0);   /* i = xE+1 <- neighbour's i = 1 */
        const bool wl = s_tab[b].wall[0] && wall_mode != FACE_NONE, wr = s_tab[b].wall[1] && wall_mode != FACE_NONE;
        const int sl = o + 2 * si, sr = o + xE * si;
        lwx[2 * q] = wl ? make_int2(o, ng ? ~sl : sl) : make_int2(-1, 0);
        lwx[2 * q + 1] = wr ? make_int2(o + (xE + 2) * si, ng ? ~sr : sr) : make_int2(-1, 0);
    }
    if (DIM == 3)
        for (int q = tid; q < nblk * Fx * Fz; q += nthr) {
            const int b = q / (Fx * Fz), w = q % (Fx * Fz), r = (w / Fz) * si + (w % Fz), o = b * vol + r;
            const int nf = s_tab[b].nbr[2], nk = s_tab[b].nbr[3];
            ly[2 * q] = (nf >= 0) ? make_int2(o, nf * vol + r + yE * sj) : make_int2(-1, 0);
            ly[2 * q + 1] = (nk >= 0) ? make_int2(o + (yE + 2) * sj, nk * vol + r + 2 * sj) : make_int2(-1, 0);
            const bool wf = s_tab[b].wall[2] && wall_mode != FACE_NONE, wb = s_tab[b].wall[3] && wall_mode != FACE_NONE;
            const int sf = o + 2 * sj, sb = o + yE * sj;
            lwy[2 * q] = wf ? make_int2(o, ng ? ~sf : sf) : make_int2(-1, 0);
            lwy[2 * q + 1] = wb ? make_int2(o + (yE + 2) * sj, ng ? ~sb : sb) : make_int2(-1, 0);
        }
    for (int q = tid; q < nblk * Fx * Fy; q += nthr) {
        const int o = q * Fz;                                       /* (b, i, j) row start */
        const int slo = o + ((z_lo_mode == FACE_PERIODIC) ? zE : 2), shi = o + ((z_hi_mode == FACE_PERIODIC) ? 2 : zE);
        lz[2 * q] = (z_lo_mode != FACE_NONE) ? make_int2(o, (z_lo_mode == FACE_NEG_MIRROR) ? ~slo : slo) : make_int2(-1, 0);
        lz[2 * q + 1] = (z_hi_mode != FACE_NONE) ? make_int2(o + zE + 2, (z_hi_mode == FACE_NEG_MIRROR) ? ~shi : shi) : make_int2(-1, 0);
    }
    __syncthreads();
    for (int q = tid; q < nblk * cnt; q += nthr) {
        const int b = q / cnt, w = q % cnt;
        const int k = w % nz, j = (w / nz) % ny, i = w / (nz * ny);
        const double d = (DIM == 3) ? 4.0 * (cx2[b * nx + i] + cy2[b * ny + j] + cz2[k]) : 4.0 * (cx2[b * nx + i] + cz2[k]);
        sD[q] = d;
        sI[q] = 1.0 / d;
    }
    const double hx = L.hx, hy = L.hy, hz = L.hz;
    const double two_h = (DIM == 3) ? 2.0 * L.hxhyhz : 2.0 * L.hzhx;
    const LapDiv dv = make_lapdiv(hx * hx, 2.0 * hx, hy * hy, 2.0 * hy, hz * hz, 2.0 * hz, DIM == 3);
    const int H = nx + ny + nz - 2;
    /* residual pass: KZ lanes (power of two >= nz) along k, the other threads over columns */
    int lkz = 0;
    while ((1 << lkz) < nz && (1 << lkz) < nthr) lkz++;
    const int KZ = 1 << lkz, kk = tid & (KZ - 1), cg = tid >> lkz, ncg = nthr >> lkz;
    __syncthreads();
#define CS_APPLY(list, n)                                                          \
    for (int e = tid; e < (n); e += nthr) {                                        \
        const int2 ds = list[e];                                                   \
        if (ds.x >= 0) sP[ds.x] = (ds.y >= 0) ? sP[ds.y] : -sP[~ds.y];             \
    }                                                                              \
    __syncthreads();
    int iter = 0;
    while (true) {
        /* ---- imposeBC of every block: x pads, y pads (with the fresh x pads), then walls x, y, z ---- */
        CS_APPLY(lx, nX)
        if (DIM == 3) { CS_APPLY(ly, nY) }
        if (wall_mode != FACE_NONE) {
            CS_APPLY(lwx, nX)
            if (DIM == 3) { CS_APPLY(lwy, nY) }
        }
        CS_APPLY(lz, nZ)
        if (iter < 0) break;

        /* ---- block Gauss-Seidel wavefront ---- */
        for (int h = 0; h < H; ++h) {
            for (int q = tid; q < ncol; q += nthr) {
                const CoarseCol cc = cols[q];
                const int k = h - cc.s;
                if ((unsigned)k < (unsigned)nz) {
                    double *c = sP + cc.off + k;
                    double t;
                    if (DIM == 3)
                        t = relax3_num(c[si], c[-si], c[sj], c[-sj], c[1], c[-1], sR[cc.off + k], cx2[cc.xi], cx1[cc.xi], hx,
                                       cy2[cc.yj], cy1[cc.yj], hy, cz2[k], cz1[k], hz, two_h);
                    else
                        t = relax2_num(c[si], c[-si], c[1], c[-1], sR[cc.off + k], cx2[cc.xi], cx1[cc.xi], hx, cz2[k], cz1[k], hz, two_h);
                    c[0] = div_inv(t, sD[cc.pt + k], sI[cc.pt + k]);
                }
            }
            __syncthreads();
        }
        double m = -1.0e-10;
        for (int q = cg; q < ncol; q += ncg) {
            const CoarseCol cc = cols[q];
            for (int k = kk; k < nz; k += KZ) {
                const double *c = sP + cc.off + k;
                const double r = sR[cc.off + k];
                double lp;
                if (DIM == 3) {
                    lp = p2 ? lap3_inv<true>(c[0], c[si], c[-si], c[sj], c[-sj], c[1], c[-1], mx2[cc.xi], mx1[cc.xi], my2[cc.yj], my1[cc.yj], mz2[k], mz1[k], dv)
                            : lap3_inv<false>(c[0], c[si], c[-si], c[sj], c[-sj], c[1], c[-1], mx2[cc.xi], mx1[cc.xi], my2[cc.yj], my1[cc.yj], mz2[k], mz1[k], dv);
                } else {
                    lp = p2 ? lap2_inv<true>(c[0], c[si], c[-si], c[1], c[-1], mx2[cc.xi], mx1[cc.xi], mz2[k], mz1[k], dv)
                            : lap2_inv<false>(c[0], c[si], c[-si], c[1], c[-1], mx2[cc.xi], mx1[cc.xi], mz2[k], mz1[k], dv);
                }
                m = fmax(m, fabs(r - lp));
            }
        }
        m = block_reduce<true>(m, sm);
        if (tid == 0) s_max = m;
        __syncthreads();
        const double gmax = s_max;
        __syncthreads();
        bool stop = gmax < tol;
        if (!stop) { iter += 1; stop = iter > maxCount; }
        if (stop) { if (tid == 0) *iters_out = iter; iter = -1; }
    }
    for (int q = tid; q < nblk * vol; q += nthr) {
        const int b = q / vol, w = q % vol;
        const int k = w % Fz - 1, j = (w / Fz) % Fy - jo, i = w / (Fz * Fy) - 1;
        Pg[b * stride + mg_at(L, i, j, k)] = sP[q];
    }
#undef CS_APPLY
#undef CS_AT
}
static inline size_t coarse_smem_bytes(const LevelView &L, int dim, int nblk) {
    const size_t Fy = (dim == 2) ? 1 : L.Fy;
    const size_t vol = (size_t)L.Fx * Fy * L.Fz, cnt = (size_t)L.nx * L.ny * L.nz;
    const size_t dbl = nblk * (2 * vol + 2 * cnt) + 4 * (size_t)nblk * (L.nx + L.ny) + 4 * (size_t)L.nz;
    const size_t nX = 2 * nblk * Fy * L.Fz, nY = (dim == 3) ? 2 * (size_t)nblk * L.Fx * L.Fz : 0, nZ = 2 * nblk * L.Fx * Fy;
    const size_t ints = 6 * (size_t)nblk * L.nx * L.ny + 2 * (2 * nX + 2 * nY + nZ);
    return sizeof(double) * (dbl + 2) + sizeof(int) * (ints + 8);
}

/* misc */
__global__ void __launch_bounds__(256) k_fill(double *__restrict__ A, const long long n, const double v) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) A[q] = v;
}
/* copy the CORE of src into dst (p:114-115); both level-0 pitched arrays */
__global__ void __launch_bounds__(256) k_copy_core(const double *__restrict__ S, double *__restrict__ Dst, const LevelView L) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int i = blockIdx.z;
    if (k >= L.nz) return;
    const long long c = mg_at(L, i, j, k);
    Dst[c] = S[c];
}
/* TEST_POISSON epilogue p:153-204: max |analytic - p| over the core */
__global__ void __launch_bounds__(256) k_analytic_dev(const double *__restrict__ P, const LevelView L, const int dim,
                                                      const int xoff, const int yoff, const int halfX, const int halfY,
                                                      double *__restrict__ part) {
    __shared__ double sm[32];
    const long long cnt = (long long)L.nx * L.ny * L.nz;
    const int ubz = L.nz - 1;
    double m = -1.0e300;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(q % L.nz);
        const int j = (int)((q / L.nz) % L.ny);
        const int i = (int)(q / ((long long)L.nz * L.ny));
        const double xDist = L.hx * (xoff + i - halfX);
        const double yDist = L.hy * (yoff + j - halfY);
        const double zDist = L.hz * (k - ubz / 2);
        const double pa = (dim == 3) ? (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0 : (xDist * xDist + zDist * zDist) / 4.0;
        m = fmax(m, fabs(pa - P[mg_at(L, i, j, k)]));
    }
    m = block_reduce<true>(m, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = m;
}
