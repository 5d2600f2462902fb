"""Configurations of the golden vectors (shared by make_golden.py and the tests)."""

GOLDEN_CASES = [
    # name, config  -- sizes kept small so every fixture is a few tens of kB
    ("d3_test_gs_np1", dict(dim=3, test_poisson=1, xInd=4, yInd=4, zInd=4, vcDepth=2, vcCount=8, gsSmooth=1, resType=2)),
    ("d3_test_jacobi_np4", dict(dim=3, test_poisson=1, xInd=4, yInd=4, zInd=4, npX=2, npY=2, vcDepth=2, vcCount=8, gsSmooth=0, resType=2)),
    ("d3_neumann_jacobi_np1", dict(dim=3, xInd=4, yInd=4, zInd=4, vcDepth=2, vcCount=10, gsSmooth=0, resType=1, mgTolerance=1e-6)),
    ("d3_neumann_tanh_gs_np2", dict(dim=3, xInd=4, yInd=3, zInd=4, npX=2, npY=1, vcDepth=2, vcCount=10, gsSmooth=1, resType=2,
                                   mgTolerance=1e-7, zGrid=2, betaZ=1.3, xGrid=2, betaX=1.1, Lx=2.0, preSmooth=3, postSmooth=1)),
    ("d3_ppp_gs_np4", dict(dim=3, xInd=4, yInd=4, zInd=4, npX=2, npY=2, vcDepth=2, vcCount=5, gsSmooth=1, resType=1,
                          xPer=1, yPer=1, zPer=1, preSmooth=4, postSmooth=4)),
    ("d3_ppn_jacobi_np8", dict(dim=3, xInd=5, yInd=4, zInd=3, npX=4, npY=2, vcDepth=1, vcCount=6, gsSmooth=0, resType=2,
                              xPer=1, yPer=1, zPer=0, zGrid=2, betaZ=1.5)),
    ("d2_ldc_gs_np4", dict(dim=2, xInd=6, zInd=6, npX=4, vcDepth=3, vcCount=7, gsSmooth=1, resType=1)),
    ("d2_test_gs_np1", dict(dim=2, test_poisson=1, xInd=5, zInd=5, vcDepth=3, vcCount=10, gsSmooth=1, resType=2)),
    ("d2_periodic_jacobi_np2", dict(dim=2, xInd=5, zInd=4, npX=2, vcDepth=2, vcCount=10, gsSmooth=0, resType=2, xPer=1, zPer=1)),
]


def inputs(o, cfg):
    """Deterministic inputs of a case: (p0 list, rhs list), one padded level-0 array per rank."""
    import numpy as np
    from tests.helpers import seeded_field, smooth_rhs
    nr = o.nranks
    _, F = o.level_dims(0)
    if cfg.get("test_poisson"):
        return [np.zeros(F) for _ in range(nr)], [np.ones(F) for _ in range(nr)]      # hydro_d3.cc:332-334: mgRHS.F = 1.0
    return [seeded_field(F, 7 + r, 0.01) for r in range(nr)], [smooth_rhs(o, r, F) + 0.3 for r in range(nr)]
