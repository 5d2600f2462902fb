/*
 * grid.h -- host-side mirror of SARAS's lib/grid/grid.h:53-319, restricted to what the multigrid path reads:
 * sizes (2^N + 1 points, x/y pencil split sharing the boundary node, grid.cc:176-215), the computational-space
 * spacings dXi/dEt/dZt, the staggered coordinates and the staggered-point metric arrays xi_x, xixx (= xi_xx),
 * xix2 (= xi_x^2) for uniform and tanh-stretched directions (grid.cc:370-519).  Member names are the
 * reference's so code written against `grid` compiles against either.  Own implementation (one routine per
 * direction instead of the reference's per-axis copies); the arithmetic expressions are the reference's, because
 * the metric values enter the stencil and parity is judged bit for bit.
 */
#ifndef GRID_H
#define GRID_H

#include <blitz/array.h>

#include "parallel.h"

class grid {
   public:
    const parser &inputParams;
    const parallel &rankData;

    blitz::TinyVector<int, 3> collocCoreSize, collocFullSize, staggrCoreSize, staggrFullSize;
    blitz::TinyVector<int, 3> padWidths, globalSize, subarrayEnds, subarrayStarts, sizeIndex;
    real dXi, dEt, dZt;
    real xLen, yLen, zLen;
    blitz::RectDomain<3> staggrCoreDomain, staggrFullDomain;

    blitz::Array<real, 1> xStaggr, yStaggr, zStaggr;
    blitz::Array<real, 1> xi_xStaggr, xixxStaggr, xix2Staggr;
    blitz::Array<real, 1> et_yStaggr, etyyStaggr, ety2Staggr;
    blitz::Array<real, 1> zt_zStaggr, ztzzStaggr, ztz2Staggr;

    grid(const parser &solParam, parallel &parallelData);

   private:
    void makeAxis(int start, int nCore, int gridType, real beta, real len, real dxi,
                  blitz::Array<real, 1> &coord, blitz::Array<real, 1> &m0, blitz::Array<real, 1> &m1, blitz::Array<real, 1> &m2);
};

#endif
