/* parser.cc -- see parser.h.  Keys and constraints follow SARAS lib/io/parser.cc:69-239 and :250-414. */
#include "parser.h"

#include <cstdlib>
#include <iostream>
#include <yaml-cpp/yaml.h>

parser::parser() {
    parseYAML("input/parameters.yaml");
    std::string why;
    if (!finalize(&why)) { std::cout << why << std::endl; exit(0); }      /* reference: message + MPI_Finalize + exit(0) */
}

parser::parser(const std::string &yamlPath) {
    parseYAML(yamlPath);
    std::string why;
    if (!finalize(&why)) { std::cout << why << std::endl; exit(0); }
}

void parser::parseYAML(const std::string &path) {
    YAML::Node y = YAML::LoadFile(path);
    YAML::Node prog = y["Program"], mesh = y["Mesh"], par = y["Parallel"], sol = y["Solver"], mg = y["Multigrid"];
    probType = prog["Problem Type"].as<int>();
    icType = prog["Initial Condition"].as<int>();
    domainType = prog["Domain Type"].as<std::string>();
    rbcType = prog["RBC Type"].as<int>();
    Re = prog["Reynolds Number"].as<real>();
    Ro = prog["Rossby Number"].as<real>();
    Ra = prog["Rayleigh Number"].as<real>();
    Pr = prog["Prandtl Number"].as<real>();
    Ta = prog["Taylor Number"].as<real>();
    Lx = prog["X Length"].as<real>();
    Ly = prog["Y Length"].as<real>();
    Lz = prog["Z Length"].as<real>();
    forceType = prog["Force"].as<int>();
    nonHgBC = prog["Heating Plate"].as<bool>();
    patchRadius = prog["Plate Radius"].as<real>();

    meshType = mesh["Mesh Type"].as<std::string>();
    betaX = mesh["X Beta"].as<real>();
    betaY = mesh["Y Beta"].as<real>();
    betaZ = mesh["Z Beta"].as<real>();
    xInd = mesh["X Index"].as<int>();
    yInd = mesh["Y Index"].as<int>();
    zInd = mesh["Z Index"].as<int>();

    nThreads = par["Number of OMP threads"].as<int>();
    npX = par["X Number of Procs"].as<int>();
    npY = par["Y Number of Procs"].as<int>();

    dScheme = sol["Differentiation Scheme"].as<int>();
    iScheme = sol["Integration Scheme"].as<int>();
    cnTolerance = sol["Solve Tolerance"].as<real>();
    restartFlag = sol["Restart Run"].as<bool>();
    useCFL = sol["Use CFL Condition"].as<bool>();
    courantNumber = sol["Courant Number"].as<real>();
    tStp = sol["Time-Step"].as<real>();
    tMax = sol["Final Time"].as<real>();
    ioCnt = sol["I/O Count"].as<int>();
    solnFormat = sol["Solution Format"].as<int>();
    fwInt = sol["Solution Write Interval"].as<real>();
    rsInt = sol["Restart Write Interval"].as<real>();
    readProbes = sol["Record Probes"].as<bool>();
    prInt = sol["Probe Time Interval"].as<real>();

    vcDepth = mg["V-Cycle Depth"].as<int>();
    vcCount = mg["V-Cycle Count"].as<int>();
    mgTolerance = mg["Solve Tolerance"].as<real>();
    gsSmooth = mg["Smoothing Method"].as<int>();
    preSmooth = mg["Pre-Smoothing Count"].as<int>();
    postSmooth = mg["Post-Smoothing Count"].as<int>();
    resType = mg["Residual Type"].as<int>();
    printResidual = mg["Print Residual"].as<bool>();
}

bool parser::finalize(std::string *why) {
    auto fail = [&](const char *m) { if (why) *why = m; return false; };
#ifdef PLANAR
    if (yInd != 0) {
        std::cout << "WARNING: Y Index parameter of YAML file is non-zero although solver has been compiled with PLANAR flag. Setting Y Index to 0" << std::endl;
        yInd = 0;
    }
#else
    if (yInd == 0) return fail("ERROR: Y Index parameter of YAML file is 0 for 3D simulation. ABORTING");
#endif
    if (npX < 1) npX = 1;
    if (npY < 1) npY = 1;
    if (domainType.length() != 3) return fail("ERROR: Domain type string is not correct. Aborting");
    if (meshType.length() != 3) return fail("ERROR: Mesh type string is not correct. Aborting");
    if (yInd == 0 && npY > 1) return fail("ERROR: More than 1 processor is specified along Y-direction, but the yInd parameter is set to 0. Aborting");
    /* local intervals per direction >= 2^(vcDepth+1)  (parser.cc:307-337) */
    const int coarsest = 1 << (vcDepth + 1);
    if ((1 << xInd) / npX < coarsest)
        return fail("ERROR: The grid size and domain decomposition along X-direction results in sub-domains too coarse to reach the V-Cycle depth specified. Aborting");
#ifndef PLANAR
    if (yInd > 0 && (1 << yInd) / npY < coarsest)
        return fail("ERROR: The grid size and domain decomposition along Y-direction results in sub-domains too coarse to reach the V-Cycle depth specified. Aborting");
#endif
    if ((1 << zInd) < coarsest) return fail("ERROR: The grid size along Z-direction is too coarse to reach the V-Cycle depth specified. Aborting");
    if (resType > 2) return fail("ERROR: The specified value for printing error at end of V-Cycles is not defined. Aborting");
    /* power-of-two decomposition (poisson.cc:370-373) */
    if ((npX & (npX - 1)) || (npY & (npY - 1))) return fail("ERROR: the number of processors along X and Y must be powers of 2 for the multigrid solver. Aborting");

    /* setGrids (parser.cc:361-391) */
    xGrid = yGrid = zGrid = 0;
    const int *dummy = nullptr; (void)dummy;
    int *g[3] = {&xGrid, &yGrid, &zGrid};
    for (int d = 0; d < 3; d++) {
        if (meshType[d] == 'S') *g[d] = 1;
        if (meshType[d] == 'D') *g[d] = 2;
    }
    /* setPeriodicity (parser.cc:406-414) */
    xPer = domainType[0] != 'N';
    yPer = domainType[1] != 'N';
    zPer = domainType[2] != 'N';
    return true;
}
