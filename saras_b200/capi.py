"""ctypes binding of include/sarasmg.h (the C-ABI of libsarasmg.so).

Plumbing only: tests and bench.py drive the CUDA path through this exactly the way the C++
drop-in (saras_b200/host/poisson.cc) does.  Loading fails loudly when the library is missing;
there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsarasmg.so")

PROC_NULL = -1
UNIQUE_ID_BYTES = 128
FLAG_TIME_RUN = 1
FLAG_PROFILE = 2
PRESSURE, RESIDUAL, TMP = 0, 1, 2
K_SMOOTH, K_RESIDUAL, K_COARSEN, K_PROLONG, K_ERROR, K_BC, K_COARSE_SOLVE, K_HALO, K_MISC, K_FUSED_RC = range(10)
KIND_NAMES = ["smooth", "residual", "coarsen", "prolong", "error", "bc", "coarse_solve", "halo", "misc", "residual+coarsen"]

_dp = C.POINTER(C.c_double)


class Desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("test_poisson", C.c_int), ("vc_depth", C.c_int), ("vc_count", C.c_int),
        ("pre_smooth", C.c_int), ("post_smooth", C.c_int), ("gs_smooth", C.c_int), ("res_type", C.c_int),
        ("print_residual", C.c_int), ("mg_tolerance", C.c_double),
        ("x_per", C.c_int), ("y_per", C.c_int), ("z_per", C.c_int),
        ("rank", C.c_int), ("n_proc", C.c_int), ("npx", C.c_int), ("npy", C.c_int), ("x_rank", C.c_int), ("y_rank", C.c_int),
        ("face_ranks", C.c_int * 4),
        ("n0", C.c_int * 3),
        ("hx", _dp), ("hy", _dp), ("hz", _dp), ("hx2", _dp), ("hz2", _dp), ("hzhx", _dp), ("hxhy", _dp), ("hyhz", _dp), ("hxhyhz", _dp),
        ("xixx", _dp), ("xix2", _dp), ("etyy", _dp), ("ety2", _dp), ("ztzz", _dp), ("ztz2", _dp),
        ("x_wall", _dp), ("y_wall", _dp), ("z_wall", _dp),
        ("device", C.c_int), ("nccl_id", C.c_void_p), ("flags", C.c_uint),
    ]


class ProjDesc(C.Structure):
    _fields_ = [("xi_x_staggr", _dp), ("et_y_staggr", _dp), ("zt_z_staggr", _dp),
                ("xi_x_colloc", _dp), ("et_y_colloc", _dp), ("zt_z_colloc", _dp),
                ("d_xi", C.c_double), ("d_et", C.c_double), ("d_zt", C.c_double)]


PROJ_VX, PROJ_VY, PROJ_VZ, PROJ_P, PROJ_PP, PROJ_RHS = range(6)
BC_NONE, BC_DIRICHLET_CC, BC_DIRICHLET_FC, BC_NEUMANN, BC_PERIODIC_CC, BC_PERIODIC_FC = range(6)


class ImplDesc(C.Structure):
    _fields_ = [("dt", C.c_double), ("Re", C.c_double), ("cn_tolerance", C.c_double),
                ("hx", C.c_double), ("hy", C.c_double), ("hz", C.c_double),
                ("hy2hz2", C.c_double), ("hz2hx2", C.c_double), ("hx2hy2", C.c_double), ("hx2hy2hz2", C.c_double),
                ("xix2_colloc", _dp), ("ety2_colloc", _dp), ("ztz2_colloc", _dp),
                ("xix2_staggr", _dp), ("ety2_staggr", _dp), ("ztz2_staggr", _dp),
                ("bc_type", (C.c_int * 6) * 3), ("bc_value", (C.c_double * 6) * 3), ("max_iterations", C.c_longlong)]


def velocity_bcs(x_per, y_per, z_per, prob_type=1):
    """hydro::initVBC (src/solvers/hydro/hydro.cc:287-371): boundary class and value per [component][wall]"""
    t = [[BC_NONE] * 6 for _ in range(3)]
    v = [[0.0] * 6 for _ in range(3)]
    if not x_per:
        t[0][0], t[0][1] = BC_DIRICHLET_FC, BC_DIRICHLET_FC
        t[1][0], t[1][1] = BC_DIRICHLET_CC, BC_DIRICHLET_CC
        t[2][0], t[2][1] = BC_DIRICHLET_CC, BC_DIRICHLET_CC
        if prob_type == 3:                  # inflow / outflow
            v[0][0] = 1.0
            t[0][1], t[1][1], t[2][1] = BC_NEUMANN, BC_NEUMANN, BC_NEUMANN
    if not y_per:
        t[0][2], t[0][3] = BC_DIRICHLET_CC, BC_DIRICHLET_CC
        t[1][2], t[1][3] = BC_DIRICHLET_FC, BC_DIRICHLET_FC
        t[2][2], t[2][3] = BC_DIRICHLET_CC, BC_DIRICHLET_CC
    if z_per:
        t[0][4] = t[0][5] = t[1][4] = t[1][5] = BC_PERIODIC_CC
        t[2][4] = t[2][5] = BC_PERIODIC_FC
    else:
        t[0][4] = t[0][5] = t[1][4] = t[1][5] = BC_DIRICHLET_CC
        t[2][4] = t[2][5] = BC_DIRICHLET_FC
        if prob_type == 1:                  # lid-driven cavity: the lid moves with Vx = 1
            v[0][5] = 1.0
    return t, v

_lib = None

SYMBOLS = [
    "sarasmg_last_error", "sarasmg_version", "sarasmg_device_count", "sarasmg_get_unique_id", "sarasmg_create", "sarasmg_destroy",
    "sarasmg_solve", "sarasmg_upload", "sarasmg_solve_resident", "sarasmg_download", "sarasmg_analytic_deviation",
    "sarasmg_level_dims", "sarasmg_level_set", "sarasmg_level_get", "sarasmg_set_level", "sarasmg_set_zero_bc",
    "sarasmg_op_update_pads", "sarasmg_op_impose_bc", "sarasmg_op_smooth", "sarasmg_op_compute_residual",
    "sarasmg_op_coarsen", "sarasmg_op_prolong", "sarasmg_op_solve_coarsest", "sarasmg_op_compute_error", "sarasmg_op_vcycle",
    "sarasmg_selftest_prolong", "sarasmg_selftest_transfer", "sarasmg_selftest_periodic",
    "sarasmg_timers", "sarasmg_kernel_launches", "sarasmg_reset_counters", "sarasmg_profile_query",
    "sarasmg_last_solve_ms", "sarasmg_synchronize", "sarasmg_debug_set_rhs_mean",
    "sarasmg_set_profile_max_level", "sarasmg_event_record", "sarasmg_event_elapsed_ms", "sarasmg_host_register",
    "sarasmg_set_vc_count", "sarasmg_host_unregister", "sarasmg_set_smoother", "sarasmg_last_rhs_mean",
    "sarasmg_impl_setup", "sarasmg_impl_upload_rhs", "sarasmg_impl_bcs", "sarasmg_impl_solve",
    "sarasmg_proj_setup", "sarasmg_proj_upload", "sarasmg_proj_download", "sarasmg_proj_divergence", "sarasmg_proj_correct", "sarasmg_proj_step",
]


def load():
    """Load libsarasmg.so (built by `python -m saras_b200.build`).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SARASMG_LIB", LIB_PATH)     # tuning: a variant built with `python -m saras_b200.build --variant NAME -D...`
    if not os.path.exists(path):
        raise RuntimeError("saras_b200: %s is missing -- run `python -m saras_b200.build` (no CPU fallback exists)" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    lib.sarasmg_last_error.restype = C.c_char_p
    lib.sarasmg_get_unique_id.argtypes = [C.c_void_p]
    lib.sarasmg_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.sarasmg_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    lib.sarasmg_destroy.argtypes = [vp]
    lib.sarasmg_solve.argtypes = [vp, _dp, _dp, ip, _dp, ip]
    lib.sarasmg_upload.argtypes = [vp, _dp, _dp]
    lib.sarasmg_solve_resident.argtypes = [vp, ip, _dp, ip]
    lib.sarasmg_download.argtypes = [vp, _dp]
    lib.sarasmg_analytic_deviation.argtypes = [vp, _dp]
    lib.sarasmg_level_dims.argtypes = [vp, C.c_int, ip, ip]
    lib.sarasmg_level_set.argtypes = [vp, C.c_int, C.c_int, _dp]
    lib.sarasmg_level_get.argtypes = [vp, C.c_int, C.c_int, _dp]
    lib.sarasmg_set_level.argtypes = [vp, C.c_int]
    lib.sarasmg_set_zero_bc.argtypes = [vp, C.c_int]
    lib.sarasmg_op_update_pads.argtypes = [vp, C.c_int]
    for f in ("sarasmg_op_impose_bc", "sarasmg_op_compute_residual", "sarasmg_op_coarsen", "sarasmg_op_prolong",
              "sarasmg_reset_counters", "sarasmg_synchronize"):
        getattr(lib, f).argtypes = [vp]
    lib.sarasmg_op_smooth.argtypes = [vp, C.c_int]
    lib.sarasmg_op_solve_coarsest.argtypes = [vp, ip]
    lib.sarasmg_op_compute_error.argtypes = [vp, C.c_int, _dp]
    lib.sarasmg_op_vcycle.argtypes = [vp, ip]
    lib.sarasmg_selftest_prolong.argtypes = [vp, _dp]
    lib.sarasmg_selftest_transfer.argtypes = [vp, _dp]
    lib.sarasmg_selftest_periodic.argtypes = [vp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
    lib.sarasmg_timers.argtypes = [vp, _dp, _dp]
    lib.sarasmg_kernel_launches.restype = C.c_longlong
    lib.sarasmg_kernel_launches.argtypes = [vp]
    lib.sarasmg_profile_query.argtypes = [vp, C.c_int, C.c_int, _dp, C.POINTER(C.c_longlong)]
    lib.sarasmg_last_solve_ms.argtypes = [vp, _dp]
    lib.sarasmg_debug_set_rhs_mean.argtypes = [vp, C.c_int, C.c_double]
    lib.sarasmg_set_profile_max_level.argtypes = [vp, C.c_int]
    lib.sarasmg_event_record.argtypes = [vp, C.c_int]
    lib.sarasmg_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, _dp]
    lib.sarasmg_host_register.argtypes = [vp, C.c_void_p, C.c_ulonglong]
    lib.sarasmg_set_vc_count.argtypes = [vp, C.c_int]
    lib.sarasmg_host_unregister.argtypes = [vp, C.c_void_p]
    lib.sarasmg_set_smoother.argtypes = [vp, C.c_int]
    lib.sarasmg_last_rhs_mean.argtypes = [vp, _dp]
    lib.sarasmg_proj_setup.argtypes = [vp, C.POINTER(ProjDesc)]
    lib.sarasmg_proj_upload.argtypes = [vp, C.c_int, _dp]
    lib.sarasmg_proj_download.argtypes = [vp, C.c_int, _dp]
    lib.sarasmg_proj_divergence.argtypes = [vp, C.c_double]
    lib.sarasmg_proj_correct.argtypes = [vp, C.c_double]
    lib.sarasmg_proj_step.argtypes = [vp, C.c_double, ip, _dp, ip]
    lib.sarasmg_impl_setup.argtypes = [vp, C.POINTER(ImplDesc)]
    lib.sarasmg_impl_upload_rhs.argtypes = [vp, C.c_int, _dp]
    lib.sarasmg_impl_bcs.argtypes = [vp, C.c_int]
    lib.sarasmg_impl_solve.argtypes = [vp, C.c_int, ip, _dp]
    _lib = lib
    return lib


class SarasMgError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Solver:
    """Thin RAII wrapper over a sarasmg_handle.

    `geom` supplies everything the C++ drop-in reads from the reference's grid/parser/parallel objects:
      n0 (3 ints), coefficient arrays hx.. (vc_depth+1 each), metric arrays (levels concatenated, padded),
      walls, rank topology.  See include/sarasmg.h.
    """

    @classmethod
    def from_handle(cls, handle, dim, vc_depth, vc_count):
        """Non-owning view of a sarasmg_handle created elsewhere (the C++ drop-in's nativeHandle())."""
        self = cls.__new__(cls)
        self.lib = load()
        self.h = C.c_void_p(handle)
        self.dim, self.depth, self.vc_count = dim, vc_depth, vc_count
        self._owned = False
        return self

    def __init__(self, *, dim, vc_depth, vc_count, pre_smooth, post_smooth, gs_smooth, res_type, mg_tolerance,
                 x_per, y_per, z_per, n0, coefs, metrics, walls=None, test_poisson=0, print_residual=0,
                 rank=0, n_proc=1, npx=1, npy=1, x_rank=0, y_rank=0, face_ranks=(PROC_NULL,) * 4,
                 device=0, nccl_id=None, flags=0):
        self.lib = load()
        d = Desc()
        d.dim, d.test_poisson, d.vc_depth, d.vc_count = dim, int(test_poisson), vc_depth, vc_count
        d.pre_smooth, d.post_smooth, d.gs_smooth, d.res_type = pre_smooth, post_smooth, int(gs_smooth), res_type
        d.print_residual, d.mg_tolerance = int(print_residual), float(mg_tolerance)
        d.x_per, d.y_per, d.z_per = int(x_per), int(y_per), int(z_per)
        d.rank, d.n_proc, d.npx, d.npy, d.x_rank, d.y_rank = rank, n_proc, npx, npy, x_rank, y_rank
        d.face_ranks = (C.c_int * 4)(*[int(f) for f in face_ranks])
        d.n0 = (C.c_int * 3)(*[int(v) for v in n0])
        self._keep = []
        for name in ("hx", "hy", "hz", "hx2", "hz2", "hzhx", "hxhy", "hyhz", "hxhyhz"):
            a = _f64(coefs[name])
            assert a.size == vc_depth + 1
            self._keep.append(a)
            setattr(d, name, _ptr(a))
        for name in ("xixx", "xix2", "etyy", "ety2", "ztzz", "ztz2"):
            a = metrics.get(name)
            if a is not None:
                a = _f64(a)
                self._keep.append(a)
            setattr(d, name, _ptr(a))
        walls = walls or {}
        for name in ("x_wall", "y_wall", "z_wall"):
            a = walls.get(name)
            if a is not None and np.size(a):
                a = _f64(a)
                self._keep.append(a)
                setattr(d, name, _ptr(a))
            else:
                setattr(d, name, None)
        d.device = device
        if nccl_id is not None:
            self._id = C.create_string_buffer(bytes(nccl_id), UNIQUE_ID_BYTES)
            d.nccl_id = C.cast(self._id, C.c_void_p)
        else:
            d.nccl_id = None
        d.flags = flags
        self.desc = d
        self.dim, self.depth, self.vc_count = dim, vc_depth, vc_count
        self.h = C.c_void_p()
        self._owned = True
        self._check(self.lib.sarasmg_create(C.byref(d), C.byref(self.h)))

    # -- plumbing --------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise SarasMgError(self.lib.sarasmg_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "_owned", True):
                self.lib.sarasmg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- mgSolve ---------------------------------------------------------------------------------
    def level_dims(self, level):
        n = (C.c_int * 3)()
        F = (C.c_int * 3)()
        self._check(self.lib.sarasmg_level_dims(self.h, level, n, F))
        return tuple(n), tuple(F)

    def _hist(self):
        return np.zeros(self.vc_count, dtype=np.float64), np.zeros(self.vc_count, dtype=np.int32), C.c_int(0)

    def mg_solve(self, p_inout, rhs):
        """poisson::mgSolve on HOST arrays (padded level-0 layout); p_inout is overwritten."""
        assert p_inout.dtype == np.float64 and p_inout.flags["C_CONTIGUOUS"] and rhs.dtype == np.float64 and rhs.flags["C_CONTIGUOUS"]
        hist, iters, cyc = self._hist()
        self._check(self.lib.sarasmg_solve(self.h, _ptr(p_inout), _ptr(rhs), C.byref(cyc), _ptr(hist),
                                           iters.ctypes.data_as(C.POINTER(C.c_int))))
        return dict(cycles=cyc.value, residuals=hist[:cyc.value].copy(), coarse_iters=iters[:cyc.value].copy())

    def upload(self, p_in, rhs):
        self._check(self.lib.sarasmg_upload(self.h, _ptr(p_in), _ptr(rhs)))

    def solve_resident(self):
        hist, iters, cyc = self._hist()
        self._check(self.lib.sarasmg_solve_resident(self.h, C.byref(cyc), _ptr(hist), iters.ctypes.data_as(C.POINTER(C.c_int))))
        return dict(cycles=cyc.value, residuals=hist[:cyc.value].copy(), coarse_iters=iters[:cyc.value].copy())

    def download(self, p_out):
        self._check(self.lib.sarasmg_download(self.h, _ptr(p_out)))

    def analytic_deviation(self):
        v = C.c_double(0.0)
        self._check(self.lib.sarasmg_analytic_deviation(self.h, C.byref(v)))
        return v.value

    # -- operators -------------------------------------------------------------------------------
    def level_set(self, level, which, a):
        a = _f64(a)
        _, F = self.level_dims(level)
        assert a.shape == tuple(F), (a.shape, F)
        self._check(self.lib.sarasmg_level_set(self.h, level, which, _ptr(a)))

    def level_get(self, level, which):
        _, F = self.level_dims(level)
        a = np.zeros(F, dtype=np.float64)
        self._check(self.lib.sarasmg_level_get(self.h, level, which, _ptr(a)))
        return a

    def set_level(self, v):
        self._check(self.lib.sarasmg_set_level(self.h, v))

    def set_zero_bc(self, z):
        self._check(self.lib.sarasmg_set_zero_bc(self.h, int(z)))

    def update_pads(self, which):
        self._check(self.lib.sarasmg_op_update_pads(self.h, which))

    def impose_bc(self):
        self._check(self.lib.sarasmg_op_impose_bc(self.h))

    def smooth(self, count):
        self._check(self.lib.sarasmg_op_smooth(self.h, count))

    def compute_residual(self):
        self._check(self.lib.sarasmg_op_compute_residual(self.h))

    def coarsen(self):
        self._check(self.lib.sarasmg_op_coarsen(self.h))

    def prolong(self):
        self._check(self.lib.sarasmg_op_prolong(self.h))

    def solve_coarsest(self):
        it = C.c_int(0)
        self._check(self.lib.sarasmg_op_solve_coarsest(self.h, C.byref(it)))
        return it.value

    def compute_error(self, norm):
        v = C.c_double(0.0)
        self._check(self.lib.sarasmg_op_compute_error(self.h, norm, C.byref(v)))
        return v.value

    def vcycle(self):
        it = C.c_int(0)
        self._check(self.lib.sarasmg_op_vcycle(self.h, C.byref(it)))
        return it.value

    def selftest_prolong(self):
        v = C.c_double(0.0)
        self._check(self.lib.sarasmg_selftest_prolong(self.h, C.byref(v)))
        return v.value

    def selftest_transfer(self):
        v = C.c_double(0.0)
        self._check(self.lib.sarasmg_selftest_transfer(self.h, C.byref(v)))
        return v.value

    def selftest_periodic(self, xs, ys, zs, lx, ly, lz):
        v = C.c_double(0.0)
        xs, zs = _f64(xs), _f64(zs)
        ys = _f64(ys) if ys is not None else None
        self._check(self.lib.sarasmg_selftest_periodic(self.h, _ptr(xs), _ptr(ys), _ptr(zs), lx, ly, lz, C.byref(v)))
        return v.value

    # -- instrumentation -------------------------------------------------------------------------
    def timers(self):
        a, b = C.c_double(0.0), C.c_double(0.0)
        self._check(self.lib.sarasmg_timers(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def kernel_launches(self):
        return int(self.lib.sarasmg_kernel_launches(self.h))

    def reset_counters(self):
        self._check(self.lib.sarasmg_reset_counters(self.h))

    def profile_query(self, kind, level):
        ms, n = C.c_double(0.0), C.c_longlong(0)
        self._check(self.lib.sarasmg_profile_query(self.h, kind, level, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_solve_ms(self):
        v = C.c_double(0.0)
        self._check(self.lib.sarasmg_last_solve_ms(self.h, C.byref(v)))
        return v.value

    def debug_set_rhs_mean(self, enable, value=0.0):
        self._check(self.lib.sarasmg_debug_set_rhs_mean(self.h, int(enable), float(value)))

    # -- device-resident projection step ---------------------------------------------
    def proj_shapes(self):
        n, _ = self.level_dims(0)
        nx, ny, nz = n
        return [(nx + 1, ny + 2, nz + 2), (nx + 2, ny + 1, nz + 2), (nx + 2, ny + 2, nz + 1), (nx + 2, ny + 2, nz + 2),
                (nx + 2, ny + 2, nz + 2), (nx + 2, ny + 2, nz + 2)]

    def proj_setup(self, staggr, colloc, spacing):
        """staggr / colloc: three 1-D arrays each (xi_x, et_y, zt_z at staggered / collocated points, first entry = index -1)"""
        d = ProjDesc()
        self._proj_keep = [_f64(a) for a in list(staggr) + list(colloc)]
        for name, a in zip(("xi_x_staggr", "et_y_staggr", "zt_z_staggr", "xi_x_colloc", "et_y_colloc", "zt_z_colloc"), self._proj_keep):
            setattr(d, name, _ptr(a))
        d.d_xi, d.d_et, d.d_zt = [float(v) for v in spacing]
        self._check(self.lib.sarasmg_proj_setup(self.h, C.byref(d)))

    def proj_upload(self, which, a):
        a = _f64(a)
        assert a.shape == self.proj_shapes()[which], (a.shape, self.proj_shapes()[which])
        self._check(self.lib.sarasmg_proj_upload(self.h, which, _ptr(a)))

    def proj_download(self, which):
        a = np.zeros(self.proj_shapes()[which], dtype=np.float64)
        self._check(self.lib.sarasmg_proj_download(self.h, which, _ptr(a)))
        return a

    def proj_divergence(self, inv_dt):
        self._check(self.lib.sarasmg_proj_divergence(self.h, float(inv_dt)))

    def proj_correct(self, dt):
        self._check(self.lib.sarasmg_proj_correct(self.h, float(dt)))

    def proj_step(self, dt):
        hist, iters, cyc = self._hist()
        self._check(self.lib.sarasmg_proj_step(self.h, float(dt), C.byref(cyc), _ptr(hist), iters.ctypes.data_as(C.POINTER(C.c_int))))
        return dict(cycles=cyc.value, residuals=hist[:cyc.value].copy(), coarse_iters=iters[:cyc.value].copy())

    # -- implicit velocity solves ---------------------------------------------------------
    def impl_setup(self, dt, Re, tol, spacing, colloc2, staggr2, bc_type, bc_value, max_iterations):
        """colloc2 / staggr2: xi_x^2, et_y^2, zt_z^2 at collocated / staggered points (first entry = index -1)"""
        d = ImplDesc()
        d.dt, d.Re, d.cn_tolerance = float(dt), float(Re), float(tol)
        hx, hy, hz = [float(v) for v in spacing]
        d.hx, d.hy, d.hz = hx, hy, hz
        d.hz2hx2 = pow(hz, 2.0) * pow(hx, 2.0)          # hydro::setCoefficients, hydro.cc:223-241 (Python's pow is C's pow)
        d.hx2hy2 = pow(hx, 2.0) * pow(hy, 2.0)
        d.hy2hz2 = pow(hy, 2.0) * pow(hz, 2.0)
        d.hx2hy2hz2 = pow(hx, 2.0) * pow(hy, 2.0) * pow(hz, 2.0)
        self._impl_keep = [_f64(a) for a in list(colloc2) + list(staggr2)]
        for name, a in zip(("xix2_colloc", "ety2_colloc", "ztz2_colloc", "xix2_staggr", "ety2_staggr", "ztz2_staggr"), self._impl_keep):
            setattr(d, name, _ptr(a))
        for c in range(3):
            for w in range(6):
                d.bc_type[c][w] = int(bc_type[c][w])
                d.bc_value[c][w] = float(bc_value[c][w])
        d.max_iterations = int(max_iterations)
        self._check(self.lib.sarasmg_impl_setup(self.h, C.byref(d)))

    def impl_upload_rhs(self, comp, a):
        a = _f64(a)
        assert a.shape == self.proj_shapes()[comp]
        self._check(self.lib.sarasmg_impl_upload_rhs(self.h, comp, _ptr(a)))

    def impl_bcs(self, comp):
        self._check(self.lib.sarasmg_impl_bcs(self.h, comp))

    def impl_solve(self, comp):
        n, e = C.c_int(0), C.c_double(0.0)
        self._check(self.lib.sarasmg_impl_solve(self.h, comp, C.byref(n), C.byref(e)))
        return n.value, e.value

    def last_rhs_mean(self):
        v = C.c_double(0.0)
        self._check(self.lib.sarasmg_last_rhs_mean(self.h, C.byref(v)))
        return v.value

    def set_vc_count(self, count):
        self._check(self.lib.sarasmg_set_vc_count(self.h, int(count)))
        self.vc_count = max(int(count), 1)

    def synchronize(self):
        self._check(self.lib.sarasmg_synchronize(self.h))

    def set_profile_max_level(self, level):
        self._check(self.lib.sarasmg_set_profile_max_level(self.h, int(level)))

    def event_record(self, slot):
        self._check(self.lib.sarasmg_event_record(self.h, int(slot)))

    def event_elapsed_ms(self, a, b):
        v = C.c_double(0.0)
        self._check(self.lib.sarasmg_event_elapsed_ms(self.h, int(a), int(b), C.byref(v)))
        return v.value

    def host_register(self, arr):
        """page-lock `arr` for this handle; the caller keeps it alive until host_unregister / close"""
        self._check(self.lib.sarasmg_host_register(self.h, C.c_void_p(arr.ctypes.data), arr.nbytes))

    def host_unregister(self, arr):
        self._check(self.lib.sarasmg_host_unregister(self.h, C.c_void_p(arr.ctypes.data)))

    def set_smoother(self, gs):
        self._check(self.lib.sarasmg_set_smoother(self.h, int(gs)))


def get_unique_id():
    lib = load()
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    if lib.sarasmg_get_unique_id(buf) != 0:
        raise SarasMgError(lib.sarasmg_last_error().decode())
    return buf.raw
