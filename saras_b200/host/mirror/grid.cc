/* grid.cc -- see grid.h */
#include "grid.h"

#include <cstdlib>
#include <iostream>

grid::grid(const parser &solParam, parallel &parallelData) : inputParams(solParam), rankData(parallelData) {
    if (inputParams.dScheme != 1) {
        if (rankData.rank == 0) std::cout << "Undefined finite differencing scheme in YAML file. ABORTING" << std::endl;
        MPI_Finalize();
        exit(0);
    }
    padWidths = 1, 1, 1;
    sizeIndex = inputParams.xInd, inputParams.yInd, inputParams.zInd;
    for (int d = 0; d < 3; d++) globalSize(d) = (sizeIndex(d) == 0) ? 1 : (1 << sizeIndex(d)) + 1;      /* grid.cc:155-164 */

    xLen = inputParams.Lx; yLen = inputParams.Ly; zLen = inputParams.Lz;
    dXi = 1.0 / real(globalSize(0) - 1);
    dEt = 1.0 / real(globalSize(1) - 1);
    dZt = 1.0 / real(globalSize(2) - 1);
#ifdef PLANAR
    yLen = 1.0;
    dEt = 1.0;
#endif

    /* grid.cc:176-215: neighbouring sub-domains share their boundary node */
    const int localNx = (globalSize(0) - 1) / rankData.npX + 1;
#ifndef PLANAR
    const int localNy = (globalSize(1) - 1) / rankData.npY + 1;
#else
    const int localNy = 1;
#endif
    const int localNz = globalSize(2);
    staggrCoreSize = localNx, localNy, localNz;
    staggrFullSize = staggrCoreSize + 2 * padWidths;
    subarrayStarts = rankData.xRank * (localNx - 1), rankData.yRank * (localNy - 1), 0;
    subarrayEnds = subarrayStarts + staggrCoreSize - 1;
    collocCoreSize = staggrCoreSize - 1;
#ifdef PLANAR
    collocCoreSize(1) += 1;
#endif
    collocFullSize = collocCoreSize + 2 * padWidths;
    staggrCoreDomain = blitz::RectDomain<3>(blitz::TinyVector<int, 3>(0, 0, 0), staggrCoreSize - 1);
    staggrFullDomain = blitz::RectDomain<3>(-padWidths, staggrCoreSize + padWidths - 1);

    makeAxis(subarrayStarts(0), localNx, inputParams.xGrid, inputParams.betaX, xLen, dXi, xStaggr, xi_xStaggr, xixxStaggr, xix2Staggr);
#ifndef PLANAR
    makeAxis(subarrayStarts(1), localNy, inputParams.yGrid, inputParams.betaY, yLen, dEt, yStaggr, et_yStaggr, etyyStaggr, ety2Staggr);
#else
    makeAxis(0, localNy, 0, 1.0, yLen, dEt, yStaggr, et_yStaggr, etyyStaggr, ety2Staggr);
    yStaggr = 1.0;      /* the y coordinates are never generated in a PLANAR build (grid.cc:308-309, 383-393) */
#endif
    makeAxis(subarrayStarts(2), localNz, inputParams.zGrid, inputParams.betaZ, zLen, dZt, zStaggr, zt_zStaggr, ztzzStaggr, ztz2Staggr);
}

/* one direction: arrays of nCore + 2 entries, index base -1 */
void grid::makeAxis(int start, int nCore, int gridType, real beta, real len, real dxi,
                    blitz::Array<real, 1> &coord, blitz::Array<real, 1> &m0, blitz::Array<real, 1> &m1, blitz::Array<real, 1> &m2) {
    const blitz::Range full(-1, nCore);
    coord.resize(full); m0.resize(full); m1.resize(full); m2.resize(full);
    /* uniform default over the whole padded range (grid.cc:370-422); xi(i) = real(global index)*dXi (grid.cc:339-351) */
    const real invLen = 1.0 / len;
    for (int i = -1; i <= nCore; i++) {
        const real xi = real(start + i) * dxi;
        coord(i) = len * xi;
        m0(i) = invLen;
        m1(i) = 0.0;
        m2(i) = pow(invLen, 2.0);
    }
    if (gridType == 2) {
        /* tangent-hyperbolic stretching on the core points only (grid.cc:458-466, 479-487, 500-508) */
        for (int i = 0; i < nCore; i++) {
            const real xi = real(start + i) * dxi;
            const real x = len * (1.0 - tanh(beta * (1.0 - 2.0 * xi)) / tanh(beta)) / 2.0;
            coord(i) = x;
            m0(i) = tanh(beta) / (beta * len * (1.0 - pow((1.0 - 2.0 * x / len) * tanh(beta), 2)));
            m1(i) = -4.0 * pow(tanh(beta), 3) * (1.0 - 2.0 * x / len) / (beta * len * len * pow(1.0 - pow(tanh(beta) * (1.0 - 2.0 * x / len), 2), 2));
            m2(i) = pow(m0(i), 2.0);
        }
    }
}
