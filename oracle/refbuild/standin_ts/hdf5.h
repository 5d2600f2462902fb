/* stand-in hdf5.h: just the handle types the reference's writer.h / reader.h / probes.h DECLARE members with, so that
 * src/solvers/hydro compiles; the HDF5-backed classes themselves (writer.cc, reader.cc, probes.cc) are not built --
 * oracle/refbuild/ts_driver.cc supplies do-nothing definitions of the few members the time stepper references. */
#ifndef SARAS_STANDIN_HDF5_H
#define SARAS_STANDIN_HDF5_H
typedef long long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;
#endif
