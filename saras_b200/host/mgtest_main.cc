/*
 * mgtest_main.cc -- the reference's Poisson system test (tests/testPoisson.sh + tests/mgTest/input/parameters.yaml)
 * with the drop-in solver: what `saras` built with -DTEST_POISSON does in its single time step
 * (hydro_d3.cc:326-343: mgRHS.F = 1.0; mgSolver.mgSolve(Pp, mgRHS)), printing the same lines
 * ("Residual after V Cycle i is ...", "Maximum absolute deviation from analytic solution is: ...").
 *
 *     mgtest [path/to/parameters.yaml]        default: input/parameters.yaml (cwd-relative, like the reference)
 *
 * MPI ranks are threads of this process (compat/mpi.h); thread r drives GPU r % device count.
 */
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "poisson.h"
#include "sarasmg.h"

static std::string g_yaml = "input/parameters.yaml";

static void rank_main(void *) {
    parser inputParams(g_yaml);
    parallel mpi(inputParams);
    grid gridData(inputParams, mpi);
#ifdef PLANAR
    multigrid_d2 mgSolver(gridData, inputParams);
#else
    multigrid_d3 mgSolver(gridData, inputParams);
#endif
    mpi.disablePeriodicAtWalls(inputParams);            /* hydro::checkPeriodic, hydro.cc:174-203 */
    char dev[16];
    int ndev = 1;
    sarasmg_device_count(&ndev);
    snprintf(dev, sizeof(dev), "%d", mpi.rank % (ndev > 0 ? ndev : 1));
    /* thread ranks share the environment, so the device is chosen by rank inside ensureDevice (rank % ndev) */
    plainsf Pp(gridData), mgRHS(gridData);
    Pp = 0.0;
#ifdef TEST_POISSON
    mgRHS = 1.0;                                        /* hydro_d3.cc:332-334 */
#else
    /* production mode has no stand-alone test in the reference; use a zero-mean cosine product */
    for (int i = 0; i < gridData.staggrCoreSize(0); i++)
        for (int j = 0; j < gridData.staggrCoreSize(1); j++)
            for (int k = 0; k < gridData.staggrCoreSize(2); k++)
                mgRHS.F(i, j, k) = cos(2.0 * M_PI * gridData.xStaggr(i) / gridData.xLen) * cos(4.0 * M_PI * gridData.zStaggr(k) / gridData.zLen);
#endif
    mgSolver.mgSolve(Pp, mgRHS);
}

int main(int argc, char **argv) {
    if (argc > 1) g_yaml = argv[1];
    unsetenv("LOCAL_RANK");
    int np = 1;
    {
        parser probe(g_yaml);
        np = probe.npX * probe.npY;
    }
    MPI_Init(&argc, &argv);
    saras_mpi_run(np, rank_main, nullptr);
    MPI_Finalize();
    return 0;
}
