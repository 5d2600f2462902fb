/* parallel.cc -- see parallel.h */
#include "parallel.h"

#include <cstdlib>
#include <iostream>

parallel::parallel(const parser &iDat) : npX(iDat.npX), npY(iDat.npY) {
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &nProc);
    if (npX * npY != nProc) {
        if (rank == 0) std::cout << "ERROR: Number of processors specified in input file does not match. Aborting" << std::endl;
        MPI_Finalize();
        exit(0);
    }
    xRank = rank % npX;
    yRank = rank / npX;

    faceRanks.resize(4);
    edgeRanks.resize(4);
    faceRanks = MPI_PROC_NULL;
    edgeRanks = MPI_PROC_NULL;
    faceRanks(0) = findRank(xRank - 1, yRank);
    faceRanks(1) = findRank(xRank + 1, yRank);
#ifndef PLANAR
    faceRanks(2) = findRank(xRank, yRank - 1);
    faceRanks(3) = findRank(xRank, yRank + 1);
    edgeRanks(0) = findRank(xRank - 1, yRank - 1);
    edgeRanks(1) = findRank(xRank - 1, yRank + 1);
    edgeRanks(2) = findRank(xRank + 1, yRank - 1);
    edgeRanks(3) = findRank(xRank + 1, yRank + 1);
#endif
    MPI_Comm_split(MPI_COMM_WORLD, yRank, xRank, &MPI_ROW_COMM);
    MPI_Comm_split(MPI_COMM_WORLD, xRank, yRank, &MPI_COL_COMM);
}

void parallel::disablePeriodicAtWalls(const parser &iDat) {
    if (!iDat.xPer) {
        if (xRank == 0) faceRanks(0) = MPI_PROC_NULL;
        if (xRank == npX - 1) faceRanks(1) = MPI_PROC_NULL;
    }
#ifdef PLANAR
    if (yRank == 0) faceRanks(2) = MPI_PROC_NULL;
    if (yRank == npY - 1) faceRanks(3) = MPI_PROC_NULL;
#else
    if (!iDat.yPer) {
        if (yRank == 0) faceRanks(2) = MPI_PROC_NULL;
        if (yRank == npY - 1) faceRanks(3) = MPI_PROC_NULL;
    }
#endif
}
