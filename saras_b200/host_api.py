"""ctypes handle over the C++ host side (saras_b200/host): the drop-in `multigrid_d2` / `multigrid_d3` classes together
with the parser -> parallel -> grid -> plainsf object model.  `HostSolver.mg_solve` is literally
`mgSolver.mgSolve(Pp, mgRHS)` (reference hydro_d3.cc:343) on numpy arrays.  Plumbing for tests and bench.py."""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}
_dp = C.POINTER(C.c_double)

DEFAULTS = dict(dim=3, test_poisson=0, xInd=5, yInd=5, zInd=5, npX=1, npY=1, xPer=0, yPer=0, zPer=0,
                xGrid=0, yGrid=0, zGrid=0, betaX=1.0, betaY=1.0, betaZ=1.0, Lx=1.0, Ly=1.0, Lz=1.0,
                vcDepth=3, vcCount=10, preSmooth=2, postSmooth=2, gsSmooth=0, resType=2,
                mgTolerance=1.0e-6, wall_null=1)


class HostParams(C.Structure):
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


BARRIER_CB = C.CFUNCTYPE(None, C.c_void_p)
ALLGATHER_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int)


def lib_path(dim, test_poisson):
    return os.path.join(_HERE, "lib", "libsarashost_d%d%s.so" % (dim, "_test" if test_poisson else ""))


def _load(dim, test_poisson):
    key = (dim, bool(test_poisson))
    if key in _LIBS:
        return _LIBS[key]
    capi.load()         # libsarasmg.so first (RTLD_GLOBAL) so the host library binds to the same copy
    path = lib_path(dim, test_poisson)
    if not os.path.exists(path):
        raise RuntimeError("saras_b200: %s is missing -- run `python -m saras_b200.build`" % path)
    lib = C.CDLL(path)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    lib.sarashost_last_error.restype = C.c_char_p
    lib.sarashost_create.restype = vp
    lib.sarashost_create.argtypes = [C.POINTER(HostParams), C.c_int]
    lib.sarashost_destroy.argtypes = [vp]
    lib.sarashost_extents.argtypes = [vp, ip]
    lib.sarashost_staggr.argtypes = [vp, C.c_int, _dp]
    lib.sarashost_rank_info.argtypes = [vp, ip]
    lib.sarashost_mgsolve.argtypes = [vp, _dp, _dp, ip, _dp, ip]
    lib.sarashost_field.restype = _dp
    lib.sarashost_field.argtypes = [vp, C.c_int]
    lib.sarashost_mgsolve_inplace.argtypes = [vp, ip]
    lib.sarashost_native.restype = vp
    lib.sarashost_native.argtypes = [vp]
    lib.sarashost_setup_array.argtypes = [vp, C.c_int, _dp]
    for f in ("sarashost_test_prolong", "sarashost_test_transfer", "sarashost_test_periodic"):
        getattr(lib, f).restype = C.c_double
        getattr(lib, f).argtypes = [vp]
    lib.sarashost_set_process_rank.argtypes = [C.c_int, C.c_int]
    lib.sarashost_set_hooks.argtypes = [BARRIER_CB, ALLGATHER_CB, vp]
    assert lib.sarashost_mode() == ((1 if dim == 2 else 0) | (2 if test_poisson else 0))
    _LIBS[key] = lib
    return lib


_hook_refs = []


def install_torch_distributed_hooks(dim, test_poisson, rank, world_size):
    """Process-rank mode (torchrun): MPI_Bcast / MPI_Barrier of the host code ride on torch.distributed (gloo or nccl)."""
    import torch
    import torch.distributed as dist
    lib = _load(dim, test_poisson)

    def _barrier(_ctx):
        dist.barrier()

    def _allgather(_ctx, send, recv, nbytes):
        mine = torch.frombuffer((C.c_ubyte * nbytes).from_address(send), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        outs = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(world_size)]
        dist.all_gather(outs, mine.to(dev))
        flat = torch.cat([o.cpu() for o in outs]).numpy()
        C.memmove(recv, flat.ctypes.data, nbytes * world_size)

    b, a = BARRIER_CB(_barrier), ALLGATHER_CB(_allgather)
    _hook_refs.extend([b, a])
    lib.sarashost_set_process_rank(rank, world_size)
    lib.sarashost_set_hooks(b, a, None)


class HostSolver:
    """parser -> parallel -> grid -> multigrid_dN -> plainsf, built from keyword parameters named like the YAML keys' members."""

    def __init__(self, print_residual=False, **kw):
        cfg = dict(DEFAULTS)
        cfg.update(kw)
        if cfg["dim"] == 2:
            cfg["yInd"] = 0
            cfg["npY"] = 1
        self.cfg = cfg
        self.lib = _load(cfg["dim"], cfg["test_poisson"])
        self.params = HostParams(**cfg)
        self.h = self.lib.sarashost_create(C.byref(self.params), int(print_residual))
        if not self.h:
            raise RuntimeError(self.lib.sarashost_last_error().decode())
        F = (C.c_int * 3)()
        self.lib.sarashost_extents(self.h, F)
        self.F = tuple(F)
        self._native = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.sarashost_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rank_info(self):
        out = (C.c_int * 6)()
        self.lib.sarashost_rank_info(self.h, out)
        return dict(xRank=out[0], yRank=out[1], faceRanks=list(out[2:6]))

    def staggr(self, d):
        out = np.zeros(self.F[d], dtype=np.float64)
        self.lib.sarashost_staggr(self.h, d, out.ctypes.data_as(_dp))
        return out

    def setup_array(self, which):
        n = self.lib.sarashost_setup_array(self.h, which, None)
        out = np.zeros(max(n, 0), dtype=np.float64)
        if n > 0:
            self.lib.sarashost_setup_array(self.h, which, out.ctypes.data_as(_dp))
        return out

    def mg_solve(self, p_inout, rhs):
        assert p_inout.shape == self.F and rhs.shape == self.F and p_inout.dtype == np.float64 and rhs.dtype == np.float64
        assert p_inout.flags["C_CONTIGUOUS"] and rhs.flags["C_CONTIGUOUS"]
        vc = max(self.cfg["vcCount"], 1)
        hist = np.zeros(vc, dtype=np.float64)
        iters = np.zeros(vc, dtype=np.int32)
        cyc = C.c_int(0)
        self.lib.sarashost_mgsolve(self.h, p_inout.ctypes.data_as(_dp), rhs.ctypes.data_as(_dp), C.byref(cyc),
                                   hist.ctypes.data_as(_dp), iters.ctypes.data_as(C.POINTER(C.c_int)))
        return dict(cycles=cyc.value, residuals=hist[:cyc.value].copy(), coarse_iters=iters[:cyc.value].copy())

    def field(self, which):
        """numpy view of the plainsf storage (0: Pp, 1: mgRHS) for the zero-copy path"""
        ptr = self.lib.sarashost_field(self.h, which)
        return np.ctypeslib.as_array(ptr, shape=self.F)

    def mg_solve_inplace(self):
        cyc = C.c_int(0)
        self.lib.sarashost_mgsolve_inplace(self.h, C.byref(cyc))
        return cyc.value

    @property
    def native(self):
        """capi.Solver view of the drop-in's device handle (creates the device hierarchy if needed)"""
        if self._native is None:
            h = self.lib.sarashost_native(self.h)
            self._native = capi.Solver.from_handle(h, self.cfg["dim"], self.cfg["vcDepth"], max(self.cfg["vcCount"], 1))
        return self._native

    def test_prolong(self):
        return self.lib.sarashost_test_prolong(self.h)

    def test_transfer(self):
        return self.lib.sarashost_test_transfer(self.h)

    def test_periodic(self):
        return self.lib.sarashost_test_periodic(self.h)
