/*
 * proj_kernels.cuh -- the two neighbours of poisson::mgSolve in the projection step of SARAS's time stepper, on the
 * device, so that Pp and mgRHS never cross PCIe:
 *
 *   before:  V.divergence(mgRHS, P);  mgRHS *= 1.0/dt;                     hydro_d3.cc:326-327, lib/field/vfield.cc:299-320
 *   after:   Pp.syncData();  P += Pp;  Pp.gradient(pressureGradient, V);
 *            pressureGradient *= dt;  V -= pressureGradient;                hydro_d3.cc:347-362, lib/field/plainsf.h:92-98,
 *                                                                           lib/parallel/mpidata.cc:278-306, sfield.cc:197-201
 *
 * Field layouts are the reference's Blitz arrays (base index -1 in every direction, z fastest, lib/field/field.cc:66-105):
 *   P, Pp, mgRHS   staggered in x, y, z     extents (nx+2, ny+2, nz+2)     -> kept in the solver's pitched level-0 layout
 *   Vx             collocated in x          extents (nx+1, ny+2, nz+2)     compact
 *   Vy             collocated in y          extents (nx+2, ny+1, nz+2)     compact
 *   Vz             collocated in z          extents (nx+2, ny+2, nz+1)     compact
 * with nx, ny, nz the staggered core sizes of the rank (2^M + 1).  Both kernels are pure streaming passes (32 B and 56 B per
 * point); expressions are associated exactly as the Blitz expressions evaluate them, -fmad=false.
 */
#pragma once
#include "mg_kernels.cuh"

struct ProjView {
    int nx, ny, nz;                     /* staggered core sizes */
    const double *xs, *ys, *zs;         /* xi_xStaggr, et_yStaggr, zt_zStaggr: pointers shifted so that [i] is index i (-1 .. n) */
    const double *xc, *yc, *zc;         /* xi_xColloc, et_yColloc, zt_zColloc: [i] is index i (-1 .. n-1) */
    double dXi, dEt, dZt;
    double rXi, rEt, rZt;               /* 1/dXi ... when p2 (the spacings are powers of two: x/d == x*(1/d) bit for bit) */
    int p2;
};

__device__ __forceinline__ long long pvx(const ProjView &V, int i, int j, int k) { return ((long long)(i + 1) * (V.ny + 2) + (j + 1)) * (V.nz + 2) + (k + 1); }
__device__ __forceinline__ long long pvy(const ProjView &V, int i, int j, int k) { return ((long long)(i + 1) * (V.ny + 1) + (j + 1)) * (V.nz + 2) + (k + 1); }
__device__ __forceinline__ long long pvz(const ProjView &V, int i, int j, int k) { return ((long long)(i + 1) * (V.ny + 2) + (j + 1)) * (V.nz + 1) + (k + 1); }
__device__ __forceinline__ double pdiv(const double a, const double d, const double r, const int p2) { return p2 ? a * r : a / d; }

/* mgRHS = 0;  mgRHS(core) = xi_x (Vx - Vx(i-1))/dXi + et_y (Vy - Vy(j-1))/dEt + zt_z (Vz - Vz(k-1))/dZt;  mgRHS *= 1/dt
 * (vfield.cc:310-319, hydro_d3.cc:327).  out = the solver's rhs input buffer (pitched level-0 layout), written over its whole padded box */
__global__ void __launch_bounds__(128) k_proj_divergence(const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
                                                         double *__restrict__ out, const LevelView L, const ProjView V, const double inv_dt) {
    /* one x-plane per blockIdx.y, its padded (j, k) rows flattened over blockIdx.x (no idle tail threads per row) */
    const int Fz = V.nz + 2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (V.ny + 2) * Fz) return;
    const int i = (int)blockIdx.y - 1, j = q / Fz - 1, k = q - (j + 1) * Fz - 1;
    double r = 0.0;
    if (i >= 0 && i < V.nx && j >= 0 && j < V.ny && k >= 0 && k < V.nz) {
        const double a = pdiv(V.xs[i] * (vx[pvx(V, i, j, k)] - vx[pvx(V, i - 1, j, k)]), V.dXi, V.rXi, V.p2);
        const double b = pdiv(V.ys[j] * (vy[pvy(V, i, j, k)] - vy[pvy(V, i, j - 1, k)]), V.dEt, V.rEt, V.p2);
        const double c = pdiv(V.zs[k] * (vz[pvz(V, i, j, k)] - vz[pvz(V, i, j, k - 1)]), V.dZt, V.rZt, V.p2);
        r = (a + b) + c;
    }
    out[mg_at(L, i, j, k)] = r * inv_dt;
}

/* Pp.syncData() on ONE rank (mpidata.cc:278-306): faces are exchanged with faceRanks (this rank itself along a periodic
 * direction, MPI_PROC_NULL after hydro::checkPeriodic otherwise), core rows only; the four edges are exchanged with
 * edgeRanks, which checkPeriodic never edits: with one rank they are always this rank.  Two launches (faces, then edges),
 * as the reference completes the faces before it starts the edges. */
__global__ void __launch_bounds__(256) k_proj_sync_faces(double *__restrict__ A, const LevelView L, const int xper, const int yper) {
    const long long nxf = (long long)L.ny * L.nz, nyf = (long long)L.nx * L.nz;
    const long long tot = (xper ? 2 * nxf : 0) + (yper ? 2 * nyf : 0);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        long long w = q;
        if (xper) {
            if (w < 2 * nxf) {
                const int hi = w >= nxf; w -= hi ? nxf : 0;
                const int j = (int)(w / L.nz), k = (int)(w % L.nz);
                /* left pad <- own plane nx-2 (what the left neighbour sends from its right side); right pad <- plane 1 */
                A[mg_at(L, hi ? L.nx : -1, j, k)] = A[mg_at(L, hi ? 1 : L.nx - 2, j, k)];
                continue;
            }
            w -= 2 * nxf;
        }
        const int hi = w >= nyf; w -= hi ? nyf : 0;
        const int i = (int)(w / L.nz), k = (int)(w % L.nz);
        A[mg_at(L, i, hi ? L.ny : -1, k)] = A[mg_at(L, i, hi ? 1 : L.ny - 2, k)];
    }
}
__global__ void __launch_bounds__(256) k_proj_sync_edges(double *__restrict__ A, const LevelView L) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 4 * L.nz; q += gridDim.x * blockDim.x) {
        const int e = q / L.nz, k = q % L.nz;
        const int xh = e >> 1, yh = e & 1;          /* pad corner (x hi?, y hi?) <- core corner of the diagonal neighbour */
        A[mg_at(L, xh ? L.nx : -1, yh ? L.ny : -1, k)] = A[mg_at(L, xh ? 1 : L.nx - 2, yh ? 1 : L.ny - 2, k)];
    }
}

/* P += Pp over the whole padded box (sfield.cc:197-201);  V -= dt * grad(Pp) on the cores of Vx, Vy, Vz (plainsf.h:92-98,
 * hydro_d3.cc:360-362; the pads of pressureGradient are zero, so the pads of V do not change) */
__global__ void __launch_bounds__(128) k_proj_correct(const double *__restrict__ pp, double *__restrict__ P, double *__restrict__ vx,
                                                      double *__restrict__ vy, double *__restrict__ vz, const LevelView L, const ProjView V,
                                                      const double dt) {
    const int Fz = V.nz + 2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (V.ny + 2) * Fz) return;
    const int i = (int)blockIdx.y - 1, j = q / Fz - 1, k = q - (j + 1) * Fz - 1;
    const long long o = mg_at(L, i, j, k);
    const double p0 = pp[o];
    P[o] = P[o] + p0;
    if (i < 0 || j < 0 || k < 0 || i >= V.nx || j >= V.ny || k >= V.nz) return;
    if (i <= V.nx - 2) {
        const double g = pdiv(V.xc[i] * (pp[o + L.plane] - p0), V.dXi, V.rXi, V.p2);
        const long long q = pvx(V, i, j, k);
        vx[q] = vx[q] - g * dt;
    }
    if (j <= V.ny - 2) {
        const double g = pdiv(V.yc[j] * (pp[o + L.pitch] - p0), V.dEt, V.rEt, V.p2);
        const long long q = pvy(V, i, j, k);
        vy[q] = vy[q] - g * dt;
    }
    if (k <= V.nz - 2) {
        const double g = pdiv(V.zc[k] * (pp[o + 1] - p0), V.dZt, V.rZt, V.p2);
        const long long q = pvz(V, i, j, k);
        vz[q] = vz[q] - g * dt;
    }
}

/* compact padded array (F_x, F_y, F_z) <-> pitched level-0 layout, for fields that live on the device in the solver's layout */
__global__ void __launch_bounds__(256) k_copy_padded(const double *__restrict__ S, double *__restrict__ D, const LevelView L) {
    const long long rows = (long long)L.Fx * L.Fy;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const long long base = row * L.pitch + MG_LEAD;
        for (int k = threadIdx.x; k < L.Fz; k += blockDim.x) D[base + k] = S[base + k];
    }
}

/* ---- Pp.syncData() between ranks (mpidata.cc:278-306) ----
 * One message per remote face (core rows only: ny x nz / nx x nz values, mpidata.cc:105-178) and one per edge (nz values of
 * the core corner column, mpidata.cc:183-259).  Sources are core points only and destinations are disjoint pads, so faces and
 * edges travel in ONE grouped ncclSend/ncclRecv.  Buffer layout (doubles): x faces [2][ny*nz], y faces [2][nx*nz], edges [4][nz];
 * side order: x lo, x hi, y lo, y hi; edges (x lo,y lo), (x lo,y hi), (x hi,y lo), (x hi,y hi). */
struct ProjSyncPlan {
    int face_remote[4];         /* exchange this face with another rank (self-periodic and wall faces are not in the messages) */
    long long off_face[4];      /* offsets of the face / edge blocks in the send and the receive buffer */
    long long off_edge[4];
};
__global__ void __launch_bounds__(256) k_proj_sync_pack(const double *__restrict__ A, const LevelView L, const ProjSyncPlan P, double *__restrict__ buf) {
    const long long nxf = (long long)L.ny * L.nz, nyf = (long long)L.nx * L.nz;
    const long long tot = 2 * nxf + 2 * nyf + 4LL * L.nz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        long long w = q;
        if (w < 2 * nxf) {
            const int hi = w >= nxf; w -= hi ? nxf : 0;
            if (!P.face_remote[hi]) continue;
            /* staggered fields share the boundary point: plane 1 goes left, plane nx-2 goes right (mpidata.cc:112, 131) */
            buf[P.off_face[hi] + w] = A[mg_at(L, hi ? L.nx - 2 : 1, (int)(w / L.nz), (int)(w % L.nz))];
            continue;
        }
        w -= 2 * nxf;
        if (w < 2 * nyf) {
            const int hi = w >= nyf; w -= hi ? nyf : 0;
            if (!P.face_remote[2 + hi]) continue;
            buf[P.off_face[2 + hi] + w] = A[mg_at(L, (int)(w / L.nz), hi ? L.ny - 2 : 1, (int)(w % L.nz))];
            continue;
        }
        w -= 2 * nyf;
        const int e = (int)(w / L.nz), k = (int)(w % L.nz);
        buf[P.off_edge[e] + k] = A[mg_at(L, (e >> 1) ? L.nx - 2 : 1, (e & 1) ? L.ny - 2 : 1, k)];
    }
}
__global__ void __launch_bounds__(256) k_proj_sync_unpack(double *__restrict__ A, const LevelView L, const ProjSyncPlan P, const double *__restrict__ buf) {
    const long long nxf = (long long)L.ny * L.nz, nyf = (long long)L.nx * L.nz;
    const long long tot = 2 * nxf + 2 * nyf + 4LL * L.nz;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        long long w = q;
        if (w < 2 * nxf) {
            const int hi = w >= nxf; w -= hi ? nxf : 0;
            if (!P.face_remote[hi]) continue;
            A[mg_at(L, hi ? L.nx : -1, (int)(w / L.nz), (int)(w % L.nz))] = buf[P.off_face[hi] + w];
            continue;
        }
        w -= 2 * nxf;
        if (w < 2 * nyf) {
            const int hi = w >= nyf; w -= hi ? nyf : 0;
            if (!P.face_remote[2 + hi]) continue;
            A[mg_at(L, (int)(w / L.nz), hi ? L.ny : -1, (int)(w % L.nz))] = buf[P.off_face[2 + hi] + w];
            continue;
        }
        w -= 2 * nyf;
        const int e = (int)(w / L.nz), k = (int)(w % L.nz);
        A[mg_at(L, (e >> 1) ? L.nx : -1, (e & 1) ? L.ny : -1, k)] = buf[P.off_edge[e] + k];
    }
}
