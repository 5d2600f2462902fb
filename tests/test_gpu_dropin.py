"""GPU: the C++ drop-in class (multigrid_d2/d3::mgSolve on plainsf fields, built by the mirror parser/grid/parallel)
against the oracle -- the call path a SARAS solver takes (hydro_d3.cc:343), end to end through the C-ABI."""
import os
import subprocess

import numpy as np
import pytest

from oracle.mgoracle import Oracle
from saras_b200 import host_api
from tests.golden.cases import inputs
from tests.helpers import max_abs, rel_linf, check_natural_mean_field

pytestmark = pytest.mark.gpu

CASES = [
    dict(dim=3, test_poisson=1, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=12, gsSmooth=1, resType=2),
    dict(dim=3, test_poisson=1, xInd=6, yInd=5, zInd=4, vcDepth=3, vcCount=8, gsSmooth=0, resType=1),
    dict(dim=3, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=20, gsSmooth=0, resType=1, mgTolerance=1e-6),
    dict(dim=3, xInd=5, yInd=4, zInd=5, vcDepth=3, vcCount=20, gsSmooth=1, resType=2, mgTolerance=1e-7, zGrid=2, betaZ=1.3, preSmooth=3, postSmooth=1),
    dict(dim=3, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=5, gsSmooth=1, resType=1, xPer=1, yPer=1, zPer=1, preSmooth=4, postSmooth=4),
    dict(dim=2, xInd=7, zInd=7, vcDepth=4, vcCount=7, gsSmooth=1, resType=1),
    dict(dim=2, test_poisson=1, xInd=6, zInd=6, vcDepth=4, vcCount=10, gsSmooth=1, resType=2),
    dict(dim=2, xInd=6, zInd=5, vcDepth=3, vcCount=10, gsSmooth=0, resType=2, xPer=1, zPer=1, zGrid=2, betaZ=1.2),
]


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: "d%d%s_%s_r%d" % (c["dim"], "t" if c.get("test_poisson") else "", "gs" if c["gsSmooth"] else "jac", c["resType"]))
def test_dropin_mgsolve_vs_oracle(cfg):
    o = Oracle(**cfg)
    p0, rhs = inputs(o, cfg)
    p_ref = p0[0].copy()
    ref = o.mg_solve([p_ref], [rhs[0].copy()])
    h = host_api.HostSolver(**cfg)
    for mode in (["natural"] if cfg.get("test_poisson") else ["injected_mean", "natural"]):
        h.native.debug_set_rhs_mean(mode == "injected_mean", o.last_global_avg())
        p = p0[0].copy()
        got = h.mg_solve(p, rhs[0].copy())
        assert got["cycles"] == ref["cycles"]
        assert np.array_equal(got["coarse_iters"], ref["coarse_iters"])
        np.testing.assert_allclose(got["residuals"], ref["residuals"], rtol=1e-10, atol=0)
        if mode == "natural" and not cfg.get("test_poisson"):
            check_natural_mean_field(o, p0, rhs, 0, p, [p_ref], "drop-in")
        else:
            assert np.array_equal(p, p_ref), max_abs(p - p_ref)
    h.close()


def test_dropin_selftests():
    """poissonTest of the reference (src/tests/units/poisson_test.cc:74-97): the three self tests, tolerance 1e-10,
    with fully periodic faceRanks as in the -DTEST_RUN harness (checkPeriodic has not run yet)"""
    for cfg in (dict(dim=3, xInd=4, yInd=4, zInd=5, vcDepth=2, xPer=1, yPer=1, zPer=1, wall_null=0),
                dict(dim=2, xInd=5, zInd=5, vcDepth=3, xPer=1, zPer=1, wall_null=0)):
        h = host_api.HostSolver(**cfg)
        assert h.test_prolong() <= 1e-10
        assert h.test_transfer() <= 1e-10
        assert h.test_periodic() <= 1e-10
        h.close()


def test_mgtest_driver_prints_reference_lines(tmp_path):
    """C++ mgTest driver (tests/testPoisson.sh equivalent): same stdout lines as the reference, values equal to the oracle's"""
    from tests.helpers import parameters_yaml
    exe = os.path.join(os.path.dirname(host_api.__file__), "lib", "mgtest_d3")
    y = tmp_path / "parameters.yaml"
    # tests/mgTest/input/parameters.yaml of the reference at Index 6, one rank: D=4 here (5 there), 10 cycles (25 there)
    y.write_text(parameters_yaml(dim=3, xInd=6, yInd=6, zInd=6, vcDepth=4, vcCount=10, gs=1, pre=2, post=2, resType=2))
    r = subprocess.run([exe, str(y)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    vals = [float(l.split(" is ")[1]) for l in r.stdout.splitlines() if l.startswith("Residual after V Cycle")]
    o = Oracle(dim=3, test_poisson=1, xInd=6, yInd=6, zInd=6, vcDepth=4, vcCount=10, gsSmooth=1, resType=2)
    _, F = o.level_dims(0)
    p = [np.zeros(F)]
    ref = o.mg_solve(p, [np.ones(F)])
    assert len(vals) == ref["cycles"]
    np.testing.assert_allclose(vals, ref["residuals"], rtol=1e-12, atol=0)       # printed with setprecision(16); the norm is a parallel reduction
    dev = [float(l.split(": ")[1]) for l in r.stdout.splitlines() if l.startswith("Maximum absolute deviation")]
    assert len(dev) == 1 and abs(dev[0] - ref["max_dev"]) <= 1e-5 * ref["max_dev"]
