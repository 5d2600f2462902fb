"""ctypes binding of oracle/_ref: the reference's own lib/poisson + lib/grid + lib/parallel + lib/io/parser sources,
compiled unmodified against stand-in headers (oracle/refbuild/Makefile).  TEST INFRASTRUCTURE ONLY.

Used (a) to pin oracle/mg_oracle.c (tests/test_oracle_vs_ref.py, tests/golden/make_golden.py) and (b) as the timed
CPU baseline of bench.py (`--impl reference`, cpu_baseline.kind == "reference").
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from .mgoracle import MgoParams, DEFAULTS

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_SRC = "/root/reference"
_LIBS = {}


def lib_name(dim, test_poisson, fast=False, dropin=False):
    if dropin:
        # the same driver and the reference's own grid / parallel / parser sources, with THIS repository's drop-in
        # poisson.{h,cc} + libsarasmg.so in place of lib/poisson (oracle/refbuild/Makefile): the product under test, not an oracle
        return "libsarasdropin_d%d%s.so" % (dim, "_test" if test_poisson else "")
    return "libsarasref_d%d%s%s.so" % (dim, "_test" if test_poisson else "", "_fast" if fast else "")


def available(dim=3, test_poisson=0, fast=False, dropin=False):
    return os.path.exists(os.path.join(REF_DIR, lib_name(dim, test_poisson, fast, dropin)))


def build():
    """Compile oracle/_ref from /root/reference (only possible where that tree exists; the GPU box uses the prebuilt files)."""
    if not os.path.isdir(REFERENCE_SRC):
        return False
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(_HERE, "refbuild")], check=True)
    return True


def _load(dim, test_poisson, fast, dropin=False):
    name = lib_name(dim, test_poisson, fast, dropin)
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        raise RuntimeError("oracle/_ref/%s missing: run `make -C oracle/refbuild` where /root/reference exists" % name)
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    lib.sarasref_mg_solve.argtypes = [C.c_char_p, C.POINTER(MgoParams), C.POINTER(dp), C.POINTER(dp), dp, dp, dp, dp, C.c_int]
    lib.sarasref_mg_solve.restype = C.c_int
    assert lib.sarasref_mode() == ((1 if dim == 2 else 0) | (2 if test_poisson else 0) | (4 if dropin else 0))
    _LIBS[name] = lib
    return lib


def _cfg(kw):
    cfg = dict(DEFAULTS)
    cfg.update(kw)
    if cfg["dim"] == 2:
        cfg["yInd"] = 0
        cfg["npY"] = 1
    return cfg


def level0_extents(cfg):
    n = [(1 << cfg["xInd"]) // cfg["npX"] + 1, 1 if cfg["dim"] == 2 else (1 << cfg["yInd"]) // cfg["npY"] + 1, (1 << cfg["zInd"]) + 1]
    return tuple(v + 2 for v in n)


def mg_solve(p_list, rhs_list, fast=False, dropin=False, **kw):
    """Run the reference's poisson::mgSolve on thread-ranks.  p_list[r] is overwritten.  Returns dict(cycles, residuals, max_dev).
    dropin=True: the reference's grid / parallel / parser objects drive this repository's drop-in solver instead (needs a GPU)."""
    cfg = _cfg(kw)
    lib = _load(cfg["dim"], cfg["test_poisson"], fast, dropin)
    nr = cfg["npX"] * cfg["npY"]
    dp = C.POINTER(C.c_double)
    for a in list(p_list) + list(rhs_list):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    P = (dp * nr)(*[a.ctypes.data_as(dp) for a in p_list])
    R = (dp * nr)(*[a.ctypes.data_as(dp) for a in rhs_list])
    hist = np.zeros(cfg["vcCount"], dtype=np.float64)
    dev = C.c_double(float("nan"))
    prm = MgoParams(**cfg)
    with tempfile.TemporaryDirectory() as wd:
        cyc = lib.sarasref_mg_solve(wd.encode(), C.byref(prm), P, R, hist.ctypes.data_as(dp), C.byref(dev), None, None, 0)
    if cyc < 0:
        raise RuntimeError("sarasref_mg_solve failed")
    return dict(cycles=cyc, residuals=hist[:cyc].copy(), max_dev=dev.value)


def setup_data(**kw):
    """Level-0 metric arrays of every rank and the per-level coefficients, as the reference's grid / poisson objects hold them."""
    cfg = _cfg(kw)
    lib = _load(cfg["dim"], cfg["test_poisson"], False)
    nr = cfg["npX"] * cfg["npY"]
    F = level0_extents(cfg)
    per_rank = 2 * (F[0] + F[1] + F[2])
    met = np.zeros(nr * per_rank, dtype=np.float64)
    coef = np.zeros(9 * (cfg["vcDepth"] + 1), dtype=np.float64)
    dp = C.POINTER(C.c_double)
    prm = MgoParams(**cfg)
    with tempfile.TemporaryDirectory() as wd:
        rc = lib.sarasref_mg_solve(wd.encode(), C.byref(prm), None, None, None, None, met.ctypes.data_as(dp), coef.ctypes.data_as(dp), 1)
    if rc < 0:
        raise RuntimeError("sarasref setup failed")
    out = []
    for r in range(nr):
        o = met[r * per_rank:(r + 1) * per_rank]
        pos = 0
        d = {}
        for q, name in enumerate(["xixx", "xix2", "etyy", "ety2", "ztzz", "ztz2"]):
            n = F[q // 2]
            d[name] = o[pos:pos + n].copy()
            pos += n
        out.append(d)
    return out, coef.reshape(cfg["vcDepth"] + 1, 9)


def proj_metrics(**kw):
    """First-derivative metrics of every rank as the reference's grid object holds them (grid.cc:370-519): per rank a dict with
    xi_xStaggr, et_yStaggr, zt_zStaggr (index -1 .. n), xi_xColloc, et_yColloc, zt_zColloc (index -1 .. n-1), (dXi, dEt, dZt)
    and the squared metrics xix2Staggr .. ztz2Colloc of the implicit solves."""
    cfg = _cfg(kw)
    lib = _load(cfg["dim"], cfg["test_poisson"], False)
    nr = cfg["npX"] * cfg["npY"]
    F = level0_extents(cfg)
    per_rank = 2 * (sum(F) + sum(f - 1 for f in F)) + 3
    buf = np.zeros(nr * per_rank, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    prm = MgoParams(**cfg)
    with tempfile.TemporaryDirectory() as wd:
        rc = lib.sarasref_mg_solve(wd.encode(), C.byref(prm), None, None, None, None, buf.ctypes.data_as(dp), None, 2)
    if rc < 0:
        raise RuntimeError("sarasref setup failed")
    out = []
    for r in range(nr):
        o = buf[r * per_rank:(r + 1) * per_rank]
        pos = 0
        d = {}
        for q, name in enumerate(["xi_xStaggr", "et_yStaggr", "zt_zStaggr", "xi_xColloc", "et_yColloc", "zt_zColloc"]):
            n = F[q % 3] - (1 if q >= 3 else 0)
            d[name] = o[pos:pos + n].copy()
            pos += n
        d["spacing"] = tuple(o[pos:pos + 3])
        pos += 3
        for q, name in enumerate(["xix2Staggr", "ety2Staggr", "ztz2Staggr", "xix2Colloc", "ety2Colloc", "ztz2Colloc"]):
            n = F[q % 3] - (1 if q >= 3 else 0)
            d[name] = o[pos:pos + n].copy()
            pos += n
        out.append(d)
    return out
