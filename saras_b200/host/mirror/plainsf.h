/*
 * plainsf.h -- host-side mirror of the one part of SARAS's lib/field/plainsf.h:51-155 the multigrid path touches:
 * the cell-centred Blitz array F with extents staggrFullSize and index base -padWidths (field.cc:72-91,
 * plainsf.cc:60-63).  In a SARAS checkout the real plainsf (constructed from an sfield) is used instead.
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
    void operator=(real a) { F = a; }
    plainsf &operator*=(real a) { F *= a; return *this; }
};

#endif
