/*
 * plainsf.h -- stand-in for /root/reference/lib/field/plainsf.h used ONLY to build oracle/_ref.
 * The real header drags in sfield/vfield/plainvf/field/mpidata/derivative (Blitz tensor-index
 * expressions that the compat header does not cover).  lib/poisson touches exactly one member of a
 * plainsf: the Blitz array F (poisson.cc:114-115, 151), whose extents and index base are those of a
 * cell-centred field: staggrFullSize with lower bound -padWidths (field.cc:72-91; plainsf.cc:60-63).
 */
#ifndef PLAINSF_H
#define PLAINSF_H

#include <blitz/array.h>
#include "grid.h"

class plainsf {
   public:
    blitz::Array<real, 3> F;
    explicit plainsf(const grid &gridData) {
        F.resize(gridData.staggrFullSize);
        F.reindexSelf(-gridData.padWidths);
        F = 0.0;
    }
};

#endif
