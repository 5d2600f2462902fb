/*
 * parallel.h -- host-side mirror of SARAS's lib/parallel/parallel.h:51-122: rank topology of the x/y pencil
 * decomposition.  Same public members and meaning: xRank = rank % npX, yRank = rank / npX (parallel.cc:89-92),
 * faceRanks = {left, right, front, back} computed periodic (parallel.cc:101-131); walls are turned into
 * MPI_PROC_NULL later by the solver (hydro::checkPeriodic, hydro.cc:174-203) -- provided here as a member
 * function so drivers without the hydro classes can apply it.  Own implementation, build scaffolding.
 */
#ifndef PARALLEL_H
#define PARALLEL_H

#include <blitz/array.h>
#include <mpi.h>

#include "parser.h"

class parallel {
   public:
    int rank, nProc;
    const int npX, npY;
    MPI_Comm MPI_ROW_COMM, MPI_COL_COMM;
    int xRank, yRank;
    blitz::Array<int, 1> faceRanks;
    blitz::Array<int, 1> edgeRanks;

    explicit parallel(const parser &iDat);

    static inline int pmod(int a, int b) { return (a % b + b) % b; }
    inline int findRank(int xR, int yR) { return pmod(yR, npY) * npX + pmod(xR, npX); }

    /* what hydro::checkPeriodic does to faceRanks (hydro.cc:183-201) */
    void disablePeriodicAtWalls(const parser &iDat);
};

#endif
