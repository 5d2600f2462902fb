"""ctypes binding of the CPU parity oracle (oracle/mg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under saras_b200/
may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

PRESSURE, RESIDUAL, TMP, SMOOTHED = 0, 1, 2, 3
XIXX, XIX2, ETYY, ETY2, ZTZZ, ZTZ2 = range(6)
HX, HY, HZ, HX2, HZ2, HZHX, HXHY, HYHZ, HXHYHZ = range(9)
COEF_NAMES = ["hx", "hy", "hz", "hx2", "hz2", "hzhx", "hxhy", "hyhz", "hxhyhz"]
METRIC_NAMES = ["xixx", "xix2", "etyy", "ety2", "ztzz", "ztz2"]


class MgoParams(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("test_poisson", C.c_int),
        ("xInd", C.c_int), ("yInd", C.c_int), ("zInd", C.c_int),
        ("npX", C.c_int), ("npY", C.c_int),
        ("xPer", C.c_int), ("yPer", C.c_int), ("zPer", C.c_int),
        ("xGrid", C.c_int), ("yGrid", C.c_int), ("zGrid", C.c_int),
        ("betaX", C.c_double), ("betaY", C.c_double), ("betaZ", C.c_double),
        ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
        ("vcDepth", C.c_int), ("vcCount", C.c_int), ("preSmooth", C.c_int),
        ("postSmooth", C.c_int), ("gsSmooth", C.c_int), ("resType", C.c_int),
        ("mgTolerance", C.c_double),
        ("wall_null", C.c_int),
    ]


def build(fast=False):
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    target = "fast" if fast else "all"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)


def _load(fast=False):
    key = "fast" if fast else "strict"
    if key in _LIBS:
        return _LIBS[key]
    name = "libmgoracle_fast.so" if fast else "libmgoracle.so"
    path = os.path.join(_HERE, "_build", name)
    srcs = [os.path.join(_HERE, f) for f in ("mg_oracle.c", "proj_oracle.c", "impl_oracle.c")]
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(f) for f in srcs):
        build(fast)
    lib = C.CDLL(path)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.mgo_create.restype = vp
    lib.mgo_create.argtypes = [C.POINTER(MgoParams)]
    lib.mgo_destroy.argtypes = [vp]
    lib.mgo_nranks.argtypes = [vp]
    lib.mgo_level_dims.argtypes = [vp, C.c_int, ip, ip]
    lib.mgo_rank_info.argtypes = [vp, C.c_int, ip]
    lib.mgo_metric.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
    lib.mgo_coef.restype = C.c_double
    lib.mgo_coef.argtypes = [vp, C.c_int, C.c_int]
    lib.mgo_grid_spacing.argtypes = [vp, dp]
    lib.mgo_staggr_coords.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.mgo_wall.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.mgo_max_count.argtypes = [vp]
    lib.mgo_array.restype = dp
    lib.mgo_array.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.mgo_set_level.argtypes = [vp, C.c_int]
    lib.mgo_get_level.argtypes = [vp]
    lib.mgo_set_zero_bc.argtypes = [vp, C.c_int]
    lib.mgo_update_pads.argtypes = [vp, C.c_int]
    for f in ("mgo_impose_bc", "mgo_compute_residual", "mgo_coarsen", "mgo_prolong", "mgo_vcycle", "mgo_lup_reset"):
        getattr(lib, f).argtypes = [vp]
    lib.mgo_smooth.argtypes = [vp, C.c_int]
    lib.mgo_solve_coarsest.argtypes = [vp]
    lib.mgo_compute_error.restype = C.c_double
    lib.mgo_compute_error.argtypes = [vp, C.c_int]
    lib.mgo_mg_solve.argtypes = [vp, C.POINTER(dp), C.POINTER(dp), dp, ip, dp]
    lib.mgo_test_prolong.restype = C.c_double
    lib.mgo_test_prolong.argtypes = [vp]
    lib.mgo_lup_count.restype = C.c_double
    lib.mgo_last_global_avg.restype = C.c_double
    lib.mgo_last_global_avg.argtypes = [vp]
    lib.mgo_set_exact_mean.argtypes = [vp, C.c_int]
    lib.projo_colloc_metric.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp]
    lib.projo_staggr_metric.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp]
    lib.projo_divergence.argtypes = [C.c_int] * 3 + [dp] * 6 + [C.c_double] * 4 + [dp]
    lib.projo_sync_one_rank.argtypes = [C.c_int] * 5 + [dp]
    lib.projo_correct.argtypes = [C.c_int] * 5 + [dp] * 8 + [C.c_double] * 4
    lib.projo_sync_multi.argtypes = [C.c_int] * 7 + [C.POINTER(dp)]
    lib.implo_metric2.argtypes = [C.c_int, dp, dp]
    lib.implo_impose_bcs.argtypes = [C.POINTER(ImploParams), C.c_int, dp]
    lib.implo_solve.argtypes = [C.POINTER(ImploParams), C.c_int, dp, dp, dp]
    lib.implo_solve.restype = C.c_int
    dpp_ = C.POINTER(dp)
    lib.implo_impose_bcs_multi.argtypes = [C.POINTER(ImploParams), C.c_int, C.c_int, C.c_int, dpp_]
    lib.implo_solve_multi.argtypes = [C.POINTER(ImploParams), C.c_int, C.c_int, C.c_int, dpp_, dpp_, dpp_, dpp_, dpp_, dpp_, dp]
    lib.implo_solve_multi.restype = C.c_int
    lib.implo_solve_T.argtypes = [C.POINTER(ImploParams), C.c_double, dp, dp, dp]
    lib.implo_solve_T.restype = C.c_int
    lib.projo_correct_nosync.argtypes = [C.c_int] * 3 + [dp] * 8 + [C.c_double] * 4
    lib.mgo_lup_count.argtypes = [vp]
    _LIBS[key] = lib
    return lib


class ImploParams(C.Structure):
    _fields_ = [("n", C.c_int * 3), ("per", C.c_int * 3), ("dt", C.c_double), ("Re", C.c_double), ("tol", C.c_double),
                ("dXi", C.c_double), ("dEt", C.c_double), ("dZt", C.c_double),
                ("m2c", C.POINTER(C.c_double) * 3), ("m2s", C.POINTER(C.c_double) * 3),
                ("bc_type", C.c_int * 6), ("bc_value", C.c_double * 6)]


DEFAULTS = dict(dim=3, test_poisson=0, xInd=5, yInd=5, zInd=5, npX=1, npY=1, xPer=0, yPer=0, zPer=0,
                xGrid=0, yGrid=0, zGrid=0, betaX=1.0, betaY=1.0, betaZ=1.0, Lx=1.0, Ly=1.0, Lz=1.0,
                vcDepth=3, vcCount=10, preSmooth=2, postSmooth=2, gsSmooth=0, resType=2,
                mgTolerance=1.0e-6, wall_null=1)


class Oracle:
    """All npX*npY ranks of the reference solver, emulated in one process."""

    def __init__(self, fast=False, **kw):
        cfg = dict(DEFAULTS)
        cfg.update(kw)
        if cfg["dim"] == 2:
            cfg["yInd"] = 0
            cfg["npY"] = 1
        self.cfg = cfg
        self.lib = _load(fast)
        self.params = MgoParams(**cfg)
        self.h = self.lib.mgo_create(C.byref(self.params))
        self.nranks = self.lib.mgo_nranks(self.h)
        self.depth = cfg["vcDepth"]

    def close(self):
        if self.h:
            self.lib.mgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- geometry ------------------------------------------------------------------------
    def level_dims(self, level):
        n = (C.c_int * 3)()
        F = (C.c_int * 3)()
        self.lib.mgo_level_dims(self.h, level, n, F)
        return tuple(n), tuple(F)

    def rank_info(self, rank):
        out = (C.c_int * 6)()
        self.lib.mgo_rank_info(self.h, rank, out)
        return dict(xRank=out[0], yRank=out[1], faceRanks=[out[2], out[3], out[4], out[5]])

    def metric(self, rank, level, which):
        _, F = self.level_dims(level)
        out = np.zeros(F[which // 2], dtype=np.float64)
        self.lib.mgo_metric(self.h, rank, level, which, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def coef(self, level, which):
        return self.lib.mgo_coef(self.h, level, which)

    def spacing(self):
        out = (C.c_double * 3)()
        self.lib.mgo_grid_spacing(self.h, out)
        return tuple(out)

    def staggr(self, rank, d):
        _, F = self.level_dims(0)
        out = np.zeros(F[d], dtype=np.float64)
        self.lib.mgo_staggr_coords(self.h, rank, d, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def wall(self, rank, which):
        _, F = self.level_dims(0)
        if self.cfg["dim"] == 3:
            shape = [(F[1], F[2]), (F[0], F[2]), (F[0], F[1])][which]
        else:
            shape = [(F[2],), (0,), (F[0],)][which]
        out = np.zeros(shape, dtype=np.float64)
        if out.size:
            self.lib.mgo_wall(self.h, rank, which, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    @property
    def max_count(self):
        return self.lib.mgo_max_count(self.h)

    # -- level arrays (numpy views onto the oracle's memory) ---------------------------------
    def array(self, rank, level, which):
        _, F = self.level_dims(level)
        ptr = self.lib.mgo_array(self.h, rank, level, which)
        return np.ctypeslib.as_array(ptr, shape=F)

    # -- operators ---------------------------------------------------------------------------
    def set_level(self, v):
        self.lib.mgo_set_level(self.h, v)

    def get_level(self):
        return self.lib.mgo_get_level(self.h)

    def set_zero_bc(self, z):
        self.lib.mgo_set_zero_bc(self.h, int(z))

    def update_pads(self, which):
        self.lib.mgo_update_pads(self.h, which)

    def impose_bc(self):
        self.lib.mgo_impose_bc(self.h)

    def smooth(self, count):
        self.lib.mgo_smooth(self.h, count)

    def compute_residual(self):
        self.lib.mgo_compute_residual(self.h)

    def coarsen(self):
        self.lib.mgo_coarsen(self.h)

    def prolong(self):
        self.lib.mgo_prolong(self.h)

    def solve_coarsest(self):
        return self.lib.mgo_solve_coarsest(self.h)

    def compute_error(self, norm):
        return self.lib.mgo_compute_error(self.h, norm)

    def vcycle(self):
        self.lib.mgo_vcycle(self.h)

    def test_prolong(self):
        return self.lib.mgo_test_prolong(self.h)

    def lups(self):
        return self.lib.mgo_lup_count(self.h)

    def set_exact_mean(self, on=True):
        """TEST MODE (not the reference's arithmetic): correctly rounded local sums in the RHS mean of mg_solve"""
        self.lib.mgo_set_exact_mean(self.h, int(on))

    def last_global_avg(self):
        return self.lib.mgo_last_global_avg(self.h)

    def lup_reset(self):
        self.lib.mgo_lup_reset(self.h)

    def mg_solve(self, p_list, rhs_list):
        """poisson::mgSolve.  p_list[r] (modified in place) and rhs_list[r]: padded level-0 arrays.
        Returns dict(cycles, residuals, coarse_iters, max_dev)."""
        nr = self.nranks
        dp = C.POINTER(C.c_double)
        for a in list(p_list) + list(rhs_list):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        P = (dp * nr)(*[a.ctypes.data_as(dp) for a in p_list])
        R = (dp * nr)(*[a.ctypes.data_as(dp) for a in rhs_list])
        vc = self.cfg["vcCount"]
        hist = np.zeros(vc, dtype=np.float64)
        iters = np.zeros(vc, dtype=np.int32)
        dev = C.c_double(0.0)
        cycles = self.lib.mgo_mg_solve(self.h, P, R, hist.ctypes.data_as(dp),
                                       iters.ctypes.data_as(C.POINTER(C.c_int)), C.byref(dev))
        return dict(cycles=cycles, residuals=hist[:cycles].copy(), coarse_iters=iters[:cycles].copy(),
                    max_dev=dev.value)


# ---------------------------------------------------------------------------------------------------------------------
# projection-step oracle (oracle/proj_oracle.c): vfield::divergence, Pp.syncData / P += Pp / V -= dt grad(Pp)
# ---------------------------------------------------------------------------------------------------------------------
def _dpp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def proj_shapes(n):
    nx, ny, nz = n
    return dict(vx=(nx + 1, ny + 2, nz + 2), vy=(nx + 2, ny + 1, nz + 2), vz=(nx + 2, ny + 2, nz + 1), sc=(nx + 2, ny + 2, nz + 2))


def proj_metrics(o, rank=0):
    """first-derivative metrics of rank `rank` of Oracle `o`: (staggered xi_x, et_y, zt_z), (collocated ...), (dXi, dEt, dZt).
    grid.cc:404-421 (uniform) and 442-519 (tanh), restated in oracle/proj_oracle.c."""
    lib = _load()
    c = o.cfg
    n0, _ = o.level_dims(0)
    info = o.rank_info(rank)
    spacing = o.spacing()
    stag, coll = [], []
    starts = [info["xRank"] * (n0[0] - 1), info["yRank"] * (n0[1] - 1), 0]
    for d, (gt, beta, ln) in enumerate(((c["xGrid"], c["betaX"], c["Lx"]), (c["yGrid"], c["betaY"], c["Ly"]), (c["zGrid"], c["betaZ"], c["Lz"]))):
        out = np.zeros(n0[d] + 1)
        lib.projo_colloc_metric(n0[d], starts[d], gt, beta, ln, spacing[d], _dpp(out))
        coll.append(out)
        s = np.zeros(n0[d] + 2)
        lib.projo_staggr_metric(n0[d], starts[d], gt, beta, ln, spacing[d], _dpp(s))
        stag.append(s)
    return stag, coll, spacing


def proj_divergence(n, vx, vy, vz, stag, spacing, inv_dt):
    rhs = np.zeros(proj_shapes(n)["sc"])
    _load().projo_divergence(n[0], n[1], n[2], _dpp(vx), _dpp(vy), _dpp(vz), _dpp(stag[0]), _dpp(stag[1]), _dpp(stag[2]),
                             spacing[0], spacing[1], spacing[2], inv_dt, _dpp(rhs))
    return rhs


def proj_correct(n, xper, yper, pp, P, vx, vy, vz, coll, spacing, dt):
    """in place on pp (syncData), P, vx, vy, vz"""
    _load().projo_correct(n[0], n[1], n[2], int(xper), int(yper), _dpp(pp), _dpp(P), _dpp(vx), _dpp(vy), _dpp(vz),
                          _dpp(coll[0]), _dpp(coll[1]), _dpp(coll[2]), spacing[0], spacing[1], spacing[2], dt)


def proj_sync_multi(o, pp_list):
    """Pp.syncData() on all ranks of Oracle `o` at once (in place on every array of pp_list, rank order)"""
    n, _ = o.level_dims(0)
    c = o.cfg
    dp = C.POINTER(C.c_double)
    ptrs = (dp * o.nranks)(*[_dpp(a) for a in pp_list])
    _load().projo_sync_multi(c["npX"], c["npY"], n[0], n[1], n[2], int(c["xPer"]), int(c["yPer"]), ptrs)


def proj_correct_nosync(n, pp, P, vx, vy, vz, coll, spacing, dt):
    """P += Pp; V -= dt grad(Pp) of one rank whose Pp is already synchronised; in place on P, vx, vy, vz"""
    _load().projo_correct_nosync(n[0], n[1], n[2], _dpp(pp), _dpp(P), _dpp(vx), _dpp(vy), _dpp(vz),
                                 _dpp(coll[0]), _dpp(coll[1]), _dpp(coll[2]), spacing[0], spacing[1], spacing[2], dt)


# ---------------------------------------------------------------------------------------------------------------------
# implicit velocity solves (oracle/impl_oracle.c)
# ---------------------------------------------------------------------------------------------------------------------
def impl_metrics2(o, rank=0):
    """xi_x^2 .. at the collocated and the staggered points (grid.cc: pow(xi_x, 2.0) element by element)"""
    stag, coll, _ = proj_metrics(o, rank)
    lib = _load()
    out = []
    for arrs in (coll, stag):
        sq = []
        for a in arrs:
            b = np.zeros_like(a)
            lib.implo_metric2(a.size, _dpp(a), _dpp(b))
            sq.append(b)
        out.append(sq)
    return out[0], out[1]


class ImplOracle:
    """hydro_d3::solveVx / Vy / Vz and imposeUBCs / VBCs / WBCs on one rank"""

    def __init__(self, o, dt, Re, tol, bc_type, bc_value):
        self.lib = _load()
        self.n, _ = o.level_dims(0)
        c = o.cfg
        self.coll2, self.stag2 = impl_metrics2(o)
        self.spacing = o.spacing()
        self.params = []
        for comp in range(3):
            p = ImploParams()
            p.n[:] = self.n
            p.per[:] = [int(c["xPer"]), int(c["yPer"]), int(c["zPer"])]
            p.dt, p.Re, p.tol = float(dt), float(Re), float(tol)
            p.dXi, p.dEt, p.dZt = self.spacing
            for d in range(3):
                p.m2c[d] = _dpp(self.coll2[d])
                p.m2s[d] = _dpp(self.stag2[d])
            p.bc_type[:] = [int(v) for v in bc_type[comp]]
            p.bc_value[:] = [float(v) for v in bc_value[comp]]
            self.params.append(p)
        self.max_iterations = (self.n[0] - 1) * (self.n[1] - 1) * (self.n[2] - 1)

    def impose_bcs(self, comp, V):
        self.lib.implo_impose_bcs(C.byref(self.params[comp]), comp, _dpp(V))

    def solve(self, comp, V, rhs):
        """in place on V; returns (sweeps, max_error); sweeps = -1 when the reference would abort"""
        e = C.c_double(0.0)
        n = self.lib.implo_solve(C.byref(self.params[comp]), comp, _dpp(V), _dpp(rhs), C.byref(e))
        return n, e.value


class ImplOracleMulti:
    """hydro_d3::solveVx / Vy / Vz and imposeUBCs / VBCs / WBCs on npX x npY ranks, all emulated in one process"""

    def __init__(self, o, dt, Re, tol, bc_type, bc_value):
        self.lib = _load()
        self.o = o
        self.n, _ = o.level_dims(0)
        c = o.cfg
        self.npx, self.npy = c["npX"], c["npY"]
        self.m2 = [impl_metrics2(o, r) for r in range(o.nranks)]          # per rank: (colloc2[3], staggr2[3])
        self.params = []
        for comp in range(3):
            p = ImploParams()
            p.n[:] = self.n
            p.per[:] = [int(c["xPer"]), int(c["yPer"]), int(c["zPer"])]
            p.dt, p.Re, p.tol = float(dt), float(Re), float(tol)
            p.dXi, p.dEt, p.dZt = o.spacing()
            for d in range(3):
                p.m2c[d] = _dpp(self.m2[0][0][d])
                p.m2s[d] = _dpp(self.m2[0][1][d])
            p.bc_type[:] = [int(v) for v in bc_type[comp]]
            p.bc_value[:] = [float(v) for v in bc_value[comp]]
            self.params.append(p)

    def _ptrs(self, arrs):
        dp = C.POINTER(C.c_double)
        return (dp * len(arrs))(*[_dpp(a) for a in arrs])

    def impose_bcs(self, comp, V_list):
        self.lib.implo_impose_bcs_multi(C.byref(self.params[comp]), comp, self.npx, self.npy, self._ptrs(V_list))

    def solve(self, comp, V_list, rhs_list):
        e = C.c_double(0.0)
        m = self.m2
        n = self.lib.implo_solve_multi(C.byref(self.params[comp]), comp, self.npx, self.npy, self._ptrs(V_list), self._ptrs(rhs_list),
                                       self._ptrs([x[0][0] for x in m]), self._ptrs([x[1][0] for x in m]),
                                       self._ptrs([x[0][1] for x in m]), self._ptrs([x[1][1] for x in m]), C.byref(e))
        return n, e.value


def temperature_bcs(x_per, y_per, z_per, prob_type=5):
    """scalar::initTBC (src/solvers/scalar/scalar.cc:186-229) without the heating-plate option: class codes as in impl_oracle.c"""
    NONE, DCC, DFC, NEU, PCC, PFC = range(6)
    t, v = [NONE] * 6, [0.0] * 6
    if not x_per:
        if prob_type == 7:
            t[0], t[1], v[0] = DCC, DCC, 1.0
        else:
            t[0], t[1] = NEU, NEU
    if not y_per:
        t[2], t[3] = NEU, NEU
    if z_per:
        t[4], t[5] = PCC, PCC
    else:
        t[4], t[5] = DCC, DCC
        if prob_type == 6:
            v[5] = 1.0
        else:
            v[4] = 1.0
    return t, v


def solve_temperature(o, dt, kappa, tol, bc_type, bc_value, T, rhs):
    """scalar_d3::solveT on one rank, in place on T; returns (sweeps, max_error)"""
    lib = _load()
    coll2, stag2 = impl_metrics2(o)
    p = ImploParams()
    n, _ = o.level_dims(0)
    c = o.cfg
    p.n[:] = n
    p.per[:] = [int(c["xPer"]), int(c["yPer"]), int(c["zPer"])]
    p.dt, p.Re, p.tol = float(dt), 1.0, float(tol)
    p.dXi, p.dEt, p.dZt = o.spacing()
    for d in range(3):
        p.m2c[d] = _dpp(coll2[d])
        p.m2s[d] = _dpp(stag2[d])
    p.bc_type[:] = [int(x) for x in bc_type]
    p.bc_value[:] = [float(x) for x in bc_value]
    e = C.c_double(0.0)
    nsw = lib.implo_solve_T(C.byref(p), float(kappa), _dpp(T), _dpp(rhs), C.byref(e))
    return nsw, e.value
