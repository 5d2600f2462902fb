"""CPU: the oracle (oracle/mg_oracle.c) against golden vectors produced by the REFERENCE ITSELF
(tests/golden/make_golden.py -> oracle/_ref, the reference's sources compiled unmodified).  This is the pin:
fields and metric arrays bit-exact, residual histories to the printed precision."""
import os

import numpy as np
import pytest

from oracle import mgoracle as mo
from oracle.mgoracle import Oracle
from tests.golden.cases import GOLDEN_CASES, inputs

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name,cfg", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_oracle_matches_reference_golden(name, cfg):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    o = Oracle(**cfg)
    p, rhs = inputs(o, cfg)
    res = o.mg_solve(p, rhs)
    assert res["cycles"] == int(g["cycles"])
    # production mode prints 17 significant digits (exact round trip); TEST_POISSON forces setprecision(16) (poisson.cc:138)
    if cfg.get("test_poisson"):
        np.testing.assert_allclose(res["residuals"], g["residuals"], rtol=1e-15, atol=0)
        assert abs(res["max_dev"] - float(g["max_dev"])) <= 1e-5 * abs(float(g["max_dev"]))      # printed with 6 digits
    else:
        assert np.array_equal(res["residuals"], g["residuals"])
    for r in range(o.nranks):
        assert np.array_equal(p[r], g["p_rank%d" % r]), "rank %d field differs from the reference" % r


@pytest.mark.parametrize("name,cfg", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_oracle_setup_matches_reference_golden(name, cfg):
    """grid metrics (grid.cc:370-519) and poisson coefficients (poisson.cc:419-438) bit-exact"""
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    o = Oracle(**cfg)
    planar = cfg["dim"] == 2
    for r in range(o.nranks):
        for q, mname in enumerate(mo.METRIC_NAMES):
            if planar and q // 2 == 1:
                continue
            ref = g["%s_rank%d" % (mname, r)]
            got = o.metric(r, 0, q)
            # the poisson copy holds the core, pads = 0 (poisson.cc:466-467); the grid array has uniform-grid values in its pads
            assert np.array_equal(got[1:-1], ref[1:-1]), (mname, r)
            # coarser levels are strided samples of level 0 (poisson.cc:467)
            for l in range(1, o.depth + 1):
                assert np.array_equal(o.metric(r, l, q)[1:-1], ref[1:-1][::2 ** l])
    coef = g["coef"]
    for l in range(o.depth + 1):
        for q, cname in enumerate(mo.COEF_NAMES):
            ref = coef[l, q]
            if ref != 0.0:      # the reference build only carries the coefficients of its own dimensionality
                assert o.coef(l, q) == ref, (cname, l)
