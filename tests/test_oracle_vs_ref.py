"""CPU: the oracle against the live reference build oracle/_ref (present wherever `make -C oracle/refbuild` ran:
this container, and the GPU box because oracle/_ref travels with the snapshot).  Wider sweep than the golden set."""
import numpy as np
import pytest

from oracle import sarasref
from oracle.mgoracle import Oracle
from tests.golden.cases import inputs

CASES = [
    dict(dim=3, test_poisson=1, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=12, gsSmooth=1, resType=2),
    dict(dim=3, test_poisson=1, xInd=5, yInd=5, zInd=5, npX=2, npY=2, vcDepth=3, vcCount=12, gsSmooth=1, resType=2),     # mgTest layout
    dict(dim=3, test_poisson=1, xInd=4, yInd=5, zInd=4, npX=1, npY=4, vcDepth=2, vcCount=6, gsSmooth=0, resType=1),
    dict(dim=3, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=20, gsSmooth=0, resType=1, mgTolerance=1e-6),
    dict(dim=3, xInd=5, yInd=4, zInd=5, vcDepth=3, vcCount=20, gsSmooth=1, resType=2, mgTolerance=1e-7, zGrid=2, betaZ=1.3, preSmooth=3, postSmooth=1),
    dict(dim=3, xInd=5, yInd=5, zInd=5, npX=2, npY=2, vcDepth=3, vcCount=5, gsSmooth=1, resType=1, xPer=1, yPer=1, zPer=1),
    dict(dim=3, xInd=5, yInd=5, zInd=5, npX=4, npY=2, vcDepth=2, vcCount=6, gsSmooth=0, resType=2, xGrid=2, yGrid=2, zGrid=2, betaX=1.2, betaY=1.1, betaZ=1.4),
    dict(dim=3, xInd=5, yInd=5, zInd=4, npX=2, npY=1, vcDepth=3, vcCount=8, gsSmooth=1, resType=2, xPer=1, yPer=0, zPer=1),
    dict(dim=3, xInd=4, yInd=4, zInd=4, vcDepth=2, vcCount=7, gsSmooth=0, resType=0),
    dict(dim=2, xInd=7, zInd=7, npX=4, vcDepth=4, vcCount=7, gsSmooth=1, resType=1),                                       # ldcTest layout
    dict(dim=2, test_poisson=1, xInd=6, zInd=6, npX=2, vcDepth=4, vcCount=10, gsSmooth=1, resType=2),
    dict(dim=2, xInd=6, zInd=5, vcDepth=3, vcCount=10, gsSmooth=0, resType=2, xPer=1, zPer=1, xGrid=2, betaX=1.2),
    # odd number of Jacobi sweeps per cycle (the outer-array swap of d3:168 leaves the levels alternating), y-only split,
    # resType 0 on several ranks (exactly one cycle), and the C1 grid at full depth on the shipped 2x2 layout
    dict(dim=3, xInd=5, yInd=4, zInd=5, vcDepth=3, vcCount=9, gsSmooth=0, resType=2, mgTolerance=1e-9, preSmooth=2, postSmooth=1),
    dict(dim=3, xInd=4, yInd=5, zInd=4, npX=1, npY=2, vcDepth=2, vcCount=8, gsSmooth=1, resType=1, zGrid=2, betaZ=1.2),
    dict(dim=3, xInd=5, yInd=5, zInd=4, npX=2, npY=2, vcDepth=2, vcCount=7, gsSmooth=0, resType=0),
    dict(dim=3, test_poisson=1, xInd=6, yInd=6, zInd=6, npX=2, npY=2, vcDepth=4, vcCount=6, gsSmooth=1, resType=2),
]


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: "d%d%s_np%dx%d_%s" % (c["dim"], "t" if c.get("test_poisson") else "", c.get("npX", 1), c.get("npY", 1), "gs" if c["gsSmooth"] else "jac"))
def test_oracle_bitexact_vs_reference(cfg):
    if not sarasref.available(cfg["dim"], cfg.get("test_poisson", 0)):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    o = Oracle(**cfg)
    p0, rhs = inputs(o, cfg)
    pa = [a.copy() for a in p0]
    ra = o.mg_solve(pa, [a.copy() for a in rhs])
    pb = [a.copy() for a in p0]
    rb = sarasref.mg_solve(pb, [a.copy() for a in rhs], **cfg)
    assert ra["cycles"] == rb["cycles"]
    if cfg.get("test_poisson"):
        np.testing.assert_allclose(ra["residuals"], rb["residuals"], rtol=1e-15, atol=0)
    elif cfg["resType"] != 0:
        assert np.array_equal(ra["residuals"], rb["residuals"])
    for r in range(o.nranks):
        assert np.array_equal(pa[r], pb[r])
