/*
 * parser.h -- host-side mirror of SARAS's lib/io/parser.h:63-123: the user parameters as public members
 * with the reference's names, filled from input/parameters.yaml (same keys, parser.cc:69-239) and checked
 * with the reference's constraints (parser.cc:250-352).  Only what the multigrid path and its tests read
 * is interpreted; everything else is carried as data.  Own implementation; build scaffolding for machines
 * without the SARAS tree -- in a SARAS checkout the real lib/io/parser.{h,cc} is used instead.
 */
#ifndef PARSER_H
#define PARSER_H

#include <math.h>
#include <string>
#include <vector>
#include <blitz/array.h>

#ifdef REAL_SINGLE
#error "the B200 multigrid path is fp64 only (north_star); build with -DREAL_DOUBLE"
#endif
#ifndef REAL_DOUBLE
#define REAL_DOUBLE
#endif
#define MPI_FP_REAL MPI_DOUBLE
#define real double

class parser {
   public:
    int ioCnt = 1, rbcType = 1, nThreads = 1;
    int npY = 1, npX = 1;
    int forceType = 0, solnFormat = 1;
    int xInd = 0, yInd = 0, zInd = 0;
    int resType = 2, vcDepth = 0, vcCount = 1;
    int gsSmooth = 0, preSmooth = 2, postSmooth = 2;
    int icType = 0, dScheme = 1, iScheme = 1, probType = 1;
    int xGrid = 0, yGrid = 0, zGrid = 0;

    bool useCFL = false, nonHgBC = false, readProbes = false, restartFlag = false, printResidual = false;
    bool xPer = false, yPer = false, zPer = false;

    real Re = 1000, Ra = 1e4, Pr = 0.71, Ta = 0, Ro = 12;
    real fwInt = 0.1, rsInt = 0.1, prInt = 0.1;
    real Lx = 1, Ly = 1, Lz = 1;
    real tStp = 1e-3, tMax = 1e-3;
    real patchRadius = 0.25, courantNumber = 0.1;
    real betaX = 1, betaY = 1, betaZ = 1;
    real cnTolerance = 1e-6, mgTolerance = 1e-6;

    std::vector<blitz::TinyVector<int, 3> > probesList;

    std::string meshType = "UUU";
    std::string domainType = "NNN";

    /* reads ./input/parameters.yaml like the reference (parser.cc:72) */
    parser();
    /* reads the given YAML file */
    explicit parser(const std::string &yamlPath);
    /* no file: the caller fills the members, then calls finalize() */
    struct no_file_t {};
    explicit parser(no_file_t) {}
    /* checkData + setGrids + setPeriodicity (parser.cc:250-414); returns false and fills `why` on a fatal configuration */
    bool finalize(std::string *why = nullptr);

   private:
    void parseYAML(const std::string &path);
};

#endif
