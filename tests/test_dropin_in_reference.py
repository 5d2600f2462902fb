"""The drop-in claim made executable: the reference's OWN grid.cc / parallel.cc / parser.cc (compiled unmodified from
/root/reference) build the objects, construct `multigrid_d3 mgSolver(gridData, inputParams)` and call
`mgSolver.mgSolve(Pp, mgRHS)` exactly as hydro_d3.cc:65,343 do -- but "poisson.h" and the library behind it are this
repository's (oracle/refbuild/Makefile, libsarasdropin_*.so, built only where /root/reference exists and shipped to the
GPU box with the snapshot).

CPU: the libraries exist, load, report the right compile-time mode, and -- there being no GPU -- the call refuses loudly.
GPU: results against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import sarasref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = [(3, 0), (3, 1), (2, 0), (2, 1)]


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.parametrize("dim,test_poisson", MODES)
def test_dropin_links_into_reference_object_model(dim, test_poisson):
    if not sarasref.available(dim, test_poisson, dropin=True):
        pytest.skip("oracle/_ref/libsarasdropin_* not built (needs /root/reference)")
    lib = sarasref._load(dim, test_poisson, False, dropin=True)
    assert lib.sarasref_mode() & 4


def test_dropin_in_reference_refuses_without_gpu():
    if not sarasref.available(3, 1, dropin=True):
        pytest.skip("oracle/_ref/libsarasdropin_* not built (needs /root/reference)")
    if _has_gpu():
        pytest.skip("a GPU is present")
    code = ("import numpy as np; from oracle import sarasref;"
            "F = sarasref.level0_extents(sarasref._cfg(dict(dim=3, test_poisson=1, xInd=4, yInd=4, zInd=4, vcDepth=2)));"
            "sarasref.mg_solve([np.zeros(F)], [np.ones(F)], dropin=True, dim=3, test_poisson=1, xInd=4, yInd=4, zInd=4, vcDepth=2, vcCount=3)")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "CUDA" in r.stderr        # no CPU fallback: message + abort, like the reference's fatal paths


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [
    dict(dim=3, test_poisson=1, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=10, gsSmooth=1, resType=2),
    dict(dim=3, xInd=5, yInd=4, zInd=5, vcDepth=3, vcCount=12, gsSmooth=0, resType=1, mgTolerance=1e-7, zGrid=2, betaZ=1.3),
    dict(dim=2, test_poisson=1, xInd=6, zInd=6, vcDepth=4, vcCount=8, gsSmooth=1, resType=2),
], ids=["d3_testpoisson_gs", "d3_neumann_tanh_jacobi", "d2_testpoisson_gs"])
def test_dropin_in_reference_matches_oracle(cfg):
    if not sarasref.available(cfg["dim"], cfg.get("test_poisson", 0), dropin=True):
        pytest.skip("oracle/_ref/libsarasdropin_* not built")
    from oracle.mgoracle import Oracle
    from tests.golden.cases import inputs
    o = Oracle(**cfg)
    p0, rhs = inputs(o, cfg)
    pa = [a.copy() for a in p0]
    ra = o.mg_solve(pa, [a.copy() for a in rhs])
    pb = [a.copy() for a in p0]
    rb = sarasref.mg_solve(pb, [a.copy() for a in rhs], dropin=True, **cfg)
    assert ra["cycles"] == rb["cycles"]
    np.testing.assert_allclose(rb["residuals"], ra["residuals"], rtol=1e-10, atol=0)
    if cfg.get("test_poisson"):
        assert np.array_equal(pa[0], pb[0])
    else:       # the device forms the correctly rounded RHS mean (helpers.check_natural_mean_field)
        from tests.helpers import check_natural_mean_field
        check_natural_mean_field(o, p0, rhs, 0, pb[0], pa, "drop-in inside the reference's object model")
