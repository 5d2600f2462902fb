/*
 * poisson.cc -- drop-in for SARAS's lib/poisson/poisson.cc + poisson_d2.cc + poisson_d3.cc (see poisson.h).
 * Host-side setup with the reference's expressions, everything else forwarded to the CUDA library.
 */
#include "poisson.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "sarasmg.h"

static void die(const char *what) {
    /* no error channel exists in mgSolve (reference: message + exit); device failures are fatal and loud */
    std::cerr << "saras-mg: " << what << ": " << sarasmg_last_error() << std::endl;
    MPI_Abort(MPI_COMM_WORLD, 1);
    abort();
}

poisson::poisson(const grid &mesh, const parser &solParam)
    : mesh(mesh), inputParams(solParam), dim(3), maxCount(0), device(nullptr), lastCycles(0) {}

multigrid_d3::multigrid_d3(const grid &mesh, const parser &solParam) : poisson(mesh, solParam) { setup(3); }
multigrid_d2::multigrid_d2(const grid &mesh, const parser &solParam) : poisson(mesh, solParam) { setup(2); }

/* the constructor sequence of poisson_d3.cc:64-87 / poisson_d2.cc:64-87 */
void poisson::setup(int dimension) {
    dim = dimension;
    strideValues.resize(inputParams.vcDepth + 1);                           /* poisson.cc:74-77 */
    for (int i = 0; i <= inputParams.vcDepth; i++) strideValues(i) = int(pow(2, i));
    setLocalSizeIndex();
    setStagBounds();
    setCoefficients();
    copyStaggrDerivs();
#ifdef TEST_POISSON
    initDirichlet();
#endif
}

/* poisson.cc:377-387 */
void poisson::setLocalSizeIndex() {
    if (dim == 2)
        localSizeIndex = blitz::TinyVector<int, 3>(mesh.sizeIndex(0) - int(log2(inputParams.npX)), mesh.sizeIndex(1), mesh.sizeIndex(2));
    else
        localSizeIndex = blitz::TinyVector<int, 3>(mesh.sizeIndex(0) - int(log2(inputParams.npX)),
                                                   mesh.sizeIndex(1) - int(log2(inputParams.npY)), mesh.sizeIndex(2));
}

/* poisson.cc:326-362: xEnd/yEnd/zEnd = 2^(localSizeIndex - level), maxCount from the coarsest padded extents */
void poisson::setStagBounds() {
    const int D = inputParams.vcDepth;
    xEnd.resize(D + 1); yEnd.resize(D + 1); zEnd.resize(D + 1);
    for (int i = 0; i <= D; i++) {
        xEnd(i) = 1 << (localSizeIndex(0) - i);
        yEnd(i) = (dim == 2) ? 0 : 1 << (localSizeIndex(1) - i);
        zEnd(i) = 1 << (localSizeIndex(2) - i);
    }
    /* cgSize = stagFull(D).ubound() - lbound() = core upper bound + 2 */
    maxCount = (xEnd(D) + 2) * (yEnd(D) + 2) * (zEnd(D) + 2);
}

/* poisson.cc:399-439, expressions kept verbatim in structure: these doubles enter every stencil evaluation */
void poisson::setCoefficients() {
    const int n = inputParams.vcDepth + 1;
    hx.resize(n); hy.resize(n); hz.resize(n); hx2.resize(n); hz2.resize(n); hzhx.resize(n); hxhy.resize(n); hyhz.resize(n); hxhyhz.resize(n);
    hy = 0.0; hx2 = 0.0; hz2 = 0.0; hxhy = 0.0; hyhz = 0.0; hxhyhz = 0.0;
    for (int i = 0; i < n; i++) {
        hx(i) = strideValues(i) * mesh.dXi;
        hz(i) = strideValues(i) * mesh.dZt;
        hzhx(i) = pow(strideValues(i), 4.0) * pow(mesh.dZt, 2.0) * pow(mesh.dXi, 2.0);
        if (dim == 2) {
            hx2(i) = pow(strideValues(i) * mesh.dXi, 2.0);
            hz2(i) = pow(strideValues(i) * mesh.dZt, 2.0);
        } else {
            hy(i) = strideValues(i) * mesh.dEt;
            hxhy(i) = pow(strideValues(i), 4.0) * pow(mesh.dXi, 2.0) * pow(mesh.dEt, 2.0);
            hyhz(i) = pow(strideValues(i), 4.0) * pow(mesh.dEt, 2.0) * pow(mesh.dZt, 2.0);
            hxhyhz(i) = pow(strideValues(i), 6.0) * pow(mesh.dXi, 2.0) * pow(mesh.dEt, 2.0) * pow(mesh.dZt, 2.0);
        }
    }
}

/* poisson.cc:453-496: per level a padded array (pads 0) whose core samples the grid's staggered metric with stride 2^l */
void poisson::copyStaggrDerivs() {
    const int D = inputParams.vcDepth;
    auto sample = [&](const blitz::Array<real, 1> &src, const blitz::Array<int, 1> &end, std::vector<double> &dst) {
        dst.clear();
        for (int n = 0; n <= D; n++) {
            dst.push_back(0.0);
            for (int i = 0; i <= end(n); i++) dst.push_back(src(i * strideValues(n)));
            dst.push_back(0.0);
        }
    };
    sample(mesh.xixxStaggr, xEnd, xixx);
    sample(mesh.xix2Staggr, xEnd, xix2);
    if (dim == 3) {
        sample(mesh.etyyStaggr, yEnd, etyy);
        sample(mesh.ety2Staggr, yEnd, ety2);
    }
    sample(mesh.ztzzStaggr, zEnd, ztzz);
    sample(mesh.ztz2Staggr, zEnd, ztz2);
}

/* poisson_d3.cc:513-572 / poisson_d2.cc:367-400: ghost-node Dirichlet values r^2/6 (3-D) or r^2/4 (2-D), level 0 */
void poisson::initDirichlet() {
    const int Fx = xEnd(0) + 3, Fy = yEnd(0) + 3, Fz = zEnd(0) + 3;
    real xDist, yDist, zDist;
    const int halfIndX = xEnd(0) * mesh.rankData.npX / 2;
    if (dim == 3) {
        const int halfIndY = yEnd(0) * mesh.rankData.npY / 2;
        xWall.assign((size_t)Fy * Fz, 0.0);
        yWall.assign((size_t)Fx * Fz, 0.0);
        zWall.assign((size_t)Fx * Fy, 0.0);
        xDist = hx(0) + mesh.inputParams.Lx / 2.0;
        for (int j = 0; j <= yEnd(0); ++j) {
            yDist = hy(0) * (mesh.rankData.yRank * yEnd(0) + j - halfIndY);
            for (int k = 0; k <= zEnd(0); ++k) {
                zDist = hz(0) * (k - zEnd(0) / 2);
                xWall[(size_t)(j + 1) * Fz + (k + 1)] = (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0;
            }
        }
        yDist = hy(0) + mesh.inputParams.Ly / 2.0;
        for (int i = 0; i <= xEnd(0); ++i) {
            xDist = hx(0) * (mesh.rankData.xRank * xEnd(0) + i - halfIndX);
            for (int k = 0; k <= zEnd(0); ++k) {
                zDist = hz(0) * (k - zEnd(0) / 2);
                yWall[(size_t)(i + 1) * Fz + (k + 1)] = (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0;
            }
        }
        zDist = hz(0) + mesh.inputParams.Lz / 2.0;
        for (int i = 0; i <= xEnd(0); ++i) {
            xDist = hx(0) * (mesh.rankData.xRank * xEnd(0) + i - halfIndX);
            for (int j = 0; j <= yEnd(0); ++j) {
                yDist = hy(0) * (mesh.rankData.yRank * yEnd(0) + j - halfIndY);
                zWall[(size_t)(i + 1) * Fy + (j + 1)] = (xDist * xDist + yDist * yDist + zDist * zDist) / 6.0;
            }
        }
    } else {
        xWall.assign((size_t)Fz, 0.0);
        zWall.assign((size_t)Fx, 0.0);
        xDist = hx(0) + mesh.inputParams.Lx / 2.0;
        for (int k = 0; k <= zEnd(0); ++k) {
            zDist = hz(0) * (k - zEnd(0) / 2);
            xWall[k + 1] = (xDist * xDist + zDist * zDist) / 4.0;
        }
        zDist = hz(0) + mesh.inputParams.Lz / 2.0;
        for (int i = 0; i <= xEnd(0); ++i) {
            xDist = hx(0) * (mesh.rankData.xRank * xEnd(0) + i - halfIndX);
            zWall[i + 1] = (xDist * xDist + zDist * zDist) / 4.0;
        }
    }
}

/* Creates the device-resident hierarchy.  Deferred to the first use so that faceRanks carries the MPI_PROC_NULL
 * entries hydro::checkPeriodic writes after this object was constructed (hydro.cc:183-201). */
void poisson::ensureDevice() {
    if (device) return;
    sarasmg_desc d;
    memset(&d, 0, sizeof(d));
    d.dim = dim;
#ifdef TEST_POISSON
    d.test_poisson = 1;
#endif
    d.vc_depth = inputParams.vcDepth;
    d.vc_count = inputParams.vcCount;
    d.pre_smooth = inputParams.preSmooth;
    d.post_smooth = inputParams.postSmooth;
    d.gs_smooth = inputParams.gsSmooth;
    d.res_type = inputParams.resType;
    d.print_residual = inputParams.printResidual ? 1 : 0;
    d.mg_tolerance = inputParams.mgTolerance;
    d.x_per = inputParams.xPer; d.y_per = inputParams.yPer; d.z_per = inputParams.zPer;
    const parallel &rd = mesh.rankData;
    d.rank = rd.rank; d.n_proc = rd.nProc; d.npx = rd.npX; d.npy = rd.npY; d.x_rank = rd.xRank; d.y_rank = rd.yRank;
    for (int f = 0; f < 4; f++) d.face_ranks[f] = (rd.faceRanks(f) == MPI_PROC_NULL) ? SARASMG_PROC_NULL : rd.faceRanks(f);
    if (dim == 2) { d.face_ranks[2] = SARASMG_PROC_NULL; d.face_ranks[3] = SARASMG_PROC_NULL; }
    d.n0[0] = xEnd(0) + 1; d.n0[1] = yEnd(0) + 1; d.n0[2] = zEnd(0) + 1;
    d.hx = hx.dataFirst(); d.hy = hy.dataFirst(); d.hz = hz.dataFirst();
    d.hx2 = hx2.dataFirst(); d.hz2 = hz2.dataFirst(); d.hzhx = hzhx.dataFirst();
    d.hxhy = hxhy.dataFirst(); d.hyhz = hyhz.dataFirst(); d.hxhyhz = hxhyhz.dataFirst();
    d.xixx = xixx.data(); d.xix2 = xix2.data();
    d.etyy = etyy.empty() ? nullptr : etyy.data(); d.ety2 = ety2.empty() ? nullptr : ety2.data();
    d.ztzz = ztzz.data(); d.ztz2 = ztz2.data();
    d.x_wall = xWall.empty() ? nullptr : xWall.data();
    d.y_wall = yWall.empty() ? nullptr : yWall.data();
    d.z_wall = zWall.empty() ? nullptr : zWall.data();

    /* one GPU per rank: SARASMG_DEVICE, else LOCAL_RANK (torchrun / mpirun wrappers), else rank modulo the device count */
    int ndev = 1;
    if (sarasmg_device_count(&ndev) != 0 || ndev < 1) die("no CUDA device");
    const char *e = getenv("SARASMG_DEVICE");
    if (!e) e = getenv("LOCAL_RANK");
    d.device = e ? atoi(e) % ndev : rd.rank % ndev;

    unsigned char id[SARASMG_UNIQUE_ID_BYTES];
    if (rd.nProc > 1) {
        /* rank 0 draws the NCCL id, everyone receives it (collective, like every call into this class) */
        if (rd.rank == 0 && sarasmg_get_unique_id(id) != 0) die("sarasmg_get_unique_id");
        MPI_Bcast(id, SARASMG_UNIQUE_ID_BYTES, MPI_BYTE, 0, MPI_COMM_WORLD);
        d.nccl_id = id;
    }
#ifdef TIME_RUN
    d.flags |= SARASMG_FLAG_TIME_RUN;
#endif
    if (getenv("SARASMG_PROFILE")) d.flags |= SARASMG_FLAG_PROFILE;
    if (sarasmg_create(&d, &device) != 0) die("sarasmg_create");
}

/* poisson.cc:100-205 */
void poisson::mgSolve(plainsf &inFn, const plainsf &rhs) {
    ensureDevice();
    const blitz::TinyVector<int, 3> want(xEnd(0) + 3, yEnd(0) + 3, zEnd(0) + 3);
    for (int dd = 0; dd < 3; dd++)
        if (inFn.F.extent(dd) != want(dd) || rhs.F.extent(dd) != want(dd)) {
            std::cerr << "saras-mg: mgSolve called with fields whose extents differ from the level-0 multigrid arrays" << std::endl;
            MPI_Abort(MPI_COMM_WORLD, 1);
        }
    /* the device layer takes the compact padded layout; Blitz fields of SARAS are contiguous, views are staged */
    blitz::Array<real, 3> pHost = inFn.F.isStorageContiguous() ? inFn.F : inFn.F.copy();
    blitz::Array<real, 3> rHost = rhs.F.isStorageContiguous() ? rhs.F : rhs.F.copy();
    /* page-lock the CALLER's field storage (SARAS allocates Pp / mgRHS once and hands the same objects in at every time
     * step; they outlive this solver, hydro.h: base-class members vs the derived-class mgSolver) -- never the temporary
     * staging copies above, whose storage dies at the end of this call.  Released in ~poisson. */
    if (!getenv("SARASMG_NO_PIN")) {
        const unsigned long long bytes = (unsigned long long)want(0) * want(1) * want(2) * sizeof(real);
        if (bytes >= (1ull << 20)) {
            if (inFn.F.isStorageContiguous()) pinField(inFn.F.dataFirst(), bytes);
            if (rhs.F.isStorageContiguous()) pinField(rhs.F.dataFirst(), bytes);
        }
    }

    lastResiduals.assign(inputParams.vcCount > 0 ? inputParams.vcCount : 1, 0.0);
    lastCoarseIters.assign(lastResiduals.size(), 0);
    if (sarasmg_solve(device, pHost.dataFirst(), rHost.dataFirst(), &lastCycles, lastResiduals.data(), lastCoarseIters.data()) != 0)
        die("sarasmg_solve");
    lastResiduals.resize(lastCycles);
    lastCoarseIters.resize(lastCycles);
    if (!inFn.F.isStorageContiguous()) inFn.F = pHost;

#ifdef TEST_POISSON
    /* poisson.cc:153-204 */
    double gloMax = 0.0;
    if (sarasmg_analytic_deviation(device, &gloMax) != 0) die("sarasmg_analytic_deviation");
    if (mesh.rankData.rank == 0) {
        std::cout << std::endl;
        std::cout << "Maximum absolute deviation from analytic solution is: " << gloMax << std::endl;
        std::cout << std::endl;
    }
#endif
}

void poisson::pinField(const void *ptr, unsigned long long bytes) {
    for (const void *q : pinnedFields) if (q == ptr) return;
    if (pinnedFields.size() >= 8) return;                   /* a solver sees two fields; never grow without bound */
    if (sarasmg_host_register(device, ptr, bytes) == 0) pinnedFields.push_back(ptr);     /* failure only costs copy speed */
}

real poisson::testProlong() {
    ensureDevice();
    double v = 0.0;
    if (sarasmg_selftest_prolong(device, &v) != 0) die("sarasmg_selftest_prolong");
    return v;
}

real poisson::testTransfer() {
    ensureDevice();
    double v = 0.0;
    if (sarasmg_selftest_transfer(device, &v) != 0) die("sarasmg_selftest_transfer");
    return v;
}

real poisson::testPeriodic() {
    ensureDevice();
    double v = 0.0;
    const double *ys = (dim == 3) ? mesh.yStaggr.dataFirst() : nullptr;
    if (sarasmg_selftest_periodic(device, mesh.xStaggr.dataFirst(), ys, mesh.zStaggr.dataFirst(), mesh.xLen, mesh.yLen, mesh.zLen, &v) != 0)
        die("sarasmg_selftest_periodic");
    return v;
}

poisson::~poisson() {
    /* under -DTIME_RUN the library prints the two "Time taken in ... within smooth" lines of poisson.cc:645-652 */
    if (device) {
        for (const void *q : pinnedFields) sarasmg_host_unregister(device, q);
        sarasmg_destroy(device);
    }
}
