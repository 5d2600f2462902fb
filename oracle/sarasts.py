"""ctypes wrapper of oracle/_ref/libsarasts_d3.so: pieces of the REFERENCE's own time stepper (hydro_d3, lib/field, lib/boundary,
mpidata; unmodified sources compiled by oracle/refbuild/Makefile) run on CPU thread-ranks.  TEST INFRASTRUCTURE: the pin of
oracle/proj_oracle.c and oracle/impl_oracle.c.  See oracle/refbuild/ts_driver.cc for the modes."""
import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libsarasts_d3.so")
DROPIN_LIB = os.path.join(_HERE, "_ref", "libsarastsdropin_d3.so")     # the same time stepper linked against THIS repository's drop-in poisson class (needs a GPU to run)
DIVERGENCE, CORRECTION, IMPOSE_BCS, IMPLICIT, STEP_TAIL, MGSOLVE, COMPUTE_DIFF, TIME_ADVANCE, SOLVE_T, COMPUTE_NLIN = range(10)
FIELDS = ("vx", "vy", "vz", "P", "Pp", "mgRHS", "rhs_x", "rhs_y", "rhs_z", "T", "rhs_T")


class TsParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("xInd", "yInd", "zInd", "npX", "npY", "xPer", "yPer", "zPer", "xGrid", "yGrid", "zGrid")] + \
               [(k, C.c_double) for k in ("betaX", "betaY", "betaZ", "Lx", "Ly", "Lz")] + \
               [(k, C.c_int) for k in ("vcDepth", "vcCount", "preSmooth", "postSmooth", "gsSmooth", "resType")] + \
               [("mgTolerance", C.c_double), ("probType", C.c_int), ("Re", C.c_double), ("dt", C.c_double), ("cnTolerance", C.c_double),
                ("icType", C.c_int), ("nSteps", C.c_int), ("outputOnly", C.c_int), ("solverKind", C.c_int), ("rbcType", C.c_int),
                ("forceType", C.c_int), ("Ra", C.c_double), ("Pr", C.c_double)]


class RankFields(C.Structure):
    _fields_ = [("f", C.POINTER(C.c_double) * 11)]


def available(dropin=False):
    return os.path.exists(DROPIN_LIB if dropin else LIB)


_libs = {}


def _load(dropin=False):
    if dropin not in _libs:
        lib = C.CDLL(DROPIN_LIB if dropin else LIB)
        lib.sarasts_extents.argtypes = [C.POINTER(TsParams), C.POINTER(C.c_int)]
        lib.sarasts_run.argtypes = [C.c_char_p, C.POINTER(TsParams), C.POINTER(RankFields), C.c_int]
        lib.sarasts_run.restype = C.c_int
        _libs[dropin] = lib
    return _libs[dropin]


def params(cfg, prob_type=1, Re=1000.0, dt=0.01, cn_tolerance=1e-6, ic_type=0, n_steps=1, output_only=(), scalar=False, rbc_type=2,
           force_type=0, Ra=1.0e4, Pr=0.71):
    """cfg: an Oracle.cfg dictionary (oracle/mgoracle.py)"""
    p = TsParams()
    for k, _ in TsParams._fields_:
        if k in cfg:
            setattr(p, k, cfg[k])
    p.probType, p.Re, p.dt, p.cnTolerance = int(prob_type), float(Re), float(dt), float(cn_tolerance)
    p.icType, p.nSteps = int(ic_type), int(n_steps)
    p.outputOnly = sum(1 << FIELDS.index(name) for name in output_only)      # these fields keep the reference's own initial values
    p.solverKind, p.rbcType, p.forceType, p.Ra, p.Pr = int(bool(scalar)), int(rbc_type), int(force_type), float(Ra), float(Pr)
    return p


def extents(p, dropin=False):
    out = (C.c_int * 33)()
    _load(dropin).sarasts_extents(C.byref(p), out)
    return {name: tuple(out[3 * q:3 * q + 3]) for q, name in enumerate(FIELDS)}


def run(p, mode, fields, dropin=False):
    """fields: list (one entry per rank) of dicts name -> C-contiguous float64 array in the field's padded layout; arrays of
    are overwritten with the reference's results (all of them, when given); missing entries keep the constructor's values.
    dropin=True: the same time stepper with this repository's drop-in poisson class (CUDA) in place of the reference's lib/poisson."""
    nr = p.npX * p.npY
    assert len(fields) == nr
    ext = extents(p, dropin)
    arr = (RankFields * nr)()
    for r, d in enumerate(fields):
        for q, name in enumerate(FIELDS):
            a = d.get(name)
            if a is None:
                arr[r].f[q] = None
                continue
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == ext[name], (name, a.shape, ext[name])
            arr[r].f[q] = a.ctypes.data_as(C.POINTER(C.c_double))
    with tempfile.TemporaryDirectory() as wd:
        rc = _load(dropin).sarasts_run(wd.encode(), C.byref(p), arr, int(mode))
    if rc != 0:
        raise RuntimeError("sarasts_run failed")
