/*
 * impl_kernels.cuh -- the implicit (Crank-Nicolson) velocity solves of SARAS's time stepper on the device:
 * hydro_d3::solveVx / solveVy / solveVz (src/solvers/hydro/hydro_d3.cc:370-533), Jacobi iterations of
 *     (1 - dt/(2 Re) Lap) V = nseRHS
 * with the velocity boundary conditions of hydro::imposeUBCs / VBCs / WBCs (hydro.cc:386-445) after every sweep and the
 * max-norm test against cnTolerance.  Same 7-point metric-weighted stencil family as the multigrid smoother; one rank.
 *
 * A component lives in its compact Blitz layout (base -1, z fastest; extent = core + 2; core = n points where the component
 * is staggered, n - 1 where it is collocated: Vx in x, Vy in y, Vz in z).  Expressions are associated exactly as the
 * reference source writes them, -fmad=false, IEEE divisions.
 */
#pragma once
#include "mg_kernels.cuh"

enum { IMPL_BC_NONE = 0, IMPL_BC_DIRICHLET_CC, IMPL_BC_DIRICHLET_FC, IMPL_BC_NEUMANN, IMPL_BC_PERIODIC_CC, IMPL_BC_PERIODIC_FC };

struct ImplGeom {
    int cs[3], ext[3], lb[3], ub[3], stag[3];       /* core size, padded extent, bulk bounds (field.cc:180-226), staggered? */
    const double *mx, *my, *mz;                     /* xi_x^2, et_y^2, zt_z^2 at this component's points, [i] = index i */
};
struct ImplCoef {
    double hy2hz2, hz2hx2, hx2hy2, hx2hy2hz2, hx, hy, hz, dt, Re, inverseRe;
    /* the loop-invariant divisors of the two point formulas and their reciprocals: x / d is rebuilt exactly from 1/d
     * (div_inv, or one multiply when d is a power of two), as in the multigrid kernels */
    double den1, rden1;         /* hx2hy2hz2 * 2.0 * Re */
    double den2, rden2;         /* Re * hx2hy2hz2 */
    double hxx, rhxx, hyy, rhyy, hzz, rhzz;     /* hx*hx, hy*hy, hz*hz */
    int p2;                     /* hx*hx, hy*hy, hz*hz are powers of two (every SARAS grid: 2^n intervals) */
};
__device__ __forceinline__ double impl_div(const double a, const double d, const double r, const int p2) { return p2 ? a * r : div_inv(a, d, r); }

__device__ __forceinline__ long long impl_at(const ImplGeom &g, int i, int j, int k) {
    return ((long long)(i + 1) * g.ext[1] + (j + 1)) * g.ext[2] + (k + 1);
}

/* guessedVelocity over the whole padded array: the Jacobi update in the bulk (hydro_d3.cc:378-384), 0 elsewhere (the
 * reference's guessedVelocity is never written outside the bulk and `V.Vx.F = guessedVelocity.Vx` copies all of it) */
__global__ void __launch_bounds__(256) k_impl_sweep(const double *__restrict__ V, const double *__restrict__ rhs, double *__restrict__ G,
                                                    const ImplGeom g, const ImplCoef c) {
    /* one x-plane per blockIdx.y, its (j, k) rows flattened over blockIdx.x: no idle tail threads when ext[2] = 2^m + 3 */
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= g.ext[1] * g.ext[2]) return;
    const int i = (int)blockIdx.y - 1, j = q / g.ext[2] - 1, k = q - (j + 1) * g.ext[2] - 1;
    const long long o = impl_at(g, i, j, k);
    double r = 0.0;
    if (i >= g.lb[0] && i <= g.ub[0] && j >= g.lb[1] && j <= g.ub[1] && k >= g.lb[2] && k <= g.ub[2]) {
        const long long sx = (long long)g.ext[1] * g.ext[2], sy = g.ext[2];
        const double mx = g.mx[i], my = g.my[j], mz = g.mz[k];
        r = (div_inv((c.hy2hz2 * mx * (V[o + sx] + V[o - sx]) +
                      c.hz2hx2 * my * (V[o + sy] + V[o - sy]) +
                      c.hx2hy2 * mz * (V[o + 1] + V[o - 1])) *
                     c.dt, c.den1, c.rden1) + rhs[o]) /
            (1.0 + div_inv(c.dt * ((c.hy2hz2 * mx + c.hz2hx2 * my + c.hx2hy2 * mz)), c.den2, c.rden2));
    }
    G[o] = r;
}

/* max over the bulk of | V - Lap(V) dt/(2 Re) - rhs | (hydro_d3.cc:393-407): per-block maxima, reduced by k_reduce_final */
__global__ void __launch_bounds__(256) k_impl_error(const double *__restrict__ V, const double *__restrict__ rhs, const ImplGeom g, const ImplCoef c,
                                                    double *__restrict__ part) {
    __shared__ double sm[32];
    const int bk = g.ub[2] - g.lb[2] + 1, bj = g.ub[1] - g.lb[1] + 1;     /* bulk rows of one x-plane, flattened */
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)blockIdx.y + g.lb[0], j = q / bk + g.lb[1], k = q - (q / bk) * bk + g.lb[2];
    double e = 0.0;
    if (q < bj * bk) {
        const long long o = impl_at(g, i, j, k);
        const long long sx = (long long)g.ext[1] * g.ext[2], sy = g.ext[2];
        const double v = V[o];
        const double lap = v - (impl_div(g.mx[i] * (V[o + sx] - 2.0 * v + V[o - sx]), c.hxx, c.rhxx, c.p2) +
                                impl_div(g.my[j] * (V[o + sy] - 2.0 * v + V[o - sy]), c.hyy, c.rhyy, c.p2) +
                                impl_div(g.mz[k] * (V[o + 1] - 2.0 * v + V[o - 1]), c.hzz, c.rhzz, c.p2)) *
                                   0.5 * c.dt * c.inverseRe;
        e = fabs(lap - rhs[o]);
    }
    e = block_reduce<true>(e, sm);
    if (threadIdx.x == 0) part[blockIdx.x + (long long)gridDim.x * blockIdx.y] = e;
}

/* syncData of one rank (mpidata.cc:278-306): periodic self faces over the core rows, then the four edge columns (always) */
__global__ void __launch_bounds__(256) k_impl_sync_faces(double *__restrict__ F, const ImplGeom g, const int xper, const int yper) {
    const long long nxf = (long long)g.cs[1] * g.cs[2], nyf = (long long)g.cs[0] * g.cs[2];
    const long long tot = (xper ? nxf : 0) + (yper ? nyf : 0);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        long long w = q;
        if (xper && w < nxf) {
            const int j = (int)(w / g.cs[2]), k = (int)(w % g.cs[2]), o = g.stag[0] ? 1 : 0, cs = g.cs[0];
            const double left = F[impl_at(g, cs - 1 - o, j, k)], right = F[impl_at(g, o, j, k)];
            F[impl_at(g, -1, j, k)] = left; F[impl_at(g, cs, j, k)] = right;
            continue;
        }
        if (xper) w -= nxf;
        const int i = (int)(w / g.cs[2]), k = (int)(w % g.cs[2]), o = g.stag[1] ? 1 : 0, cs = g.cs[1];
        const double front = F[impl_at(g, i, cs - 1 - o, k)], back = F[impl_at(g, i, o, k)];
        F[impl_at(g, i, -1, k)] = front; F[impl_at(g, i, cs, k)] = back;
    }
}
__global__ void __launch_bounds__(256) k_impl_sync_edges(double *__restrict__ F, const ImplGeom g) {
    const int ox = g.stag[0] ? 1 : 0, oy = g.stag[1] ? 1 : 0, cx = g.cs[0], cy = g.cs[1];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < g.cs[2]; k += gridDim.x * blockDim.x) {
        const double c00 = F[impl_at(g, cx - 1 - ox, cy - 1 - oy, k)], c01 = F[impl_at(g, cx - 1 - ox, oy, k)];
        const double c10 = F[impl_at(g, ox, cy - 1 - oy, k)], c11 = F[impl_at(g, ox, oy, k)];
        F[impl_at(g, -1, -1, k)] = c00; F[impl_at(g, -1, cy, k)] = c01; F[impl_at(g, cx, -1, k)] = c10; F[impl_at(g, cx, cy, k)] = c11;
    }
}
/* one boundary object (lib/boundary/*.cc:71-84): plane `dst` of axis `ax` over the WHOLE extent of the other two axes */
__global__ void __launch_bounds__(256) k_impl_plane(double *__restrict__ F, const ImplGeom g, const int ax, const int dst, const int src,
                                                    const int mode, const double v) {
    const int a1 = (ax == 0) ? 1 : 0, a2 = (ax == 2) ? 1 : 2;
    const long long tot = (long long)g.ext[a1] * g.ext[a2];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        int d[3], s[3];
        d[a1] = s[a1] = (int)(q / g.ext[a2]) - 1;
        d[a2] = s[a2] = (int)(q % g.ext[a2]) - 1;
        d[ax] = dst; s[ax] = src;
        const double f = F[impl_at(g, s[0], s[1], s[2])];
        double r;
        if (mode == IMPL_BC_DIRICHLET_CC) r = v;
        else if (mode == IMPL_BC_DIRICHLET_FC) r = 2.0 * v - f;
        else r = f;
        F[impl_at(g, d[0], d[1], d[2])] = r;
    }
}
