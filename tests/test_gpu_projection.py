"""Device-resident projection step: vfield::divergence -> mgSolve -> Pp.syncData / P += Pp / V -= dt grad(Pp)
through the C-ABI (sarasmg_proj_*) against oracle/proj_oracle.c + the multigrid oracle on identical seeded fields.
Bar: bit-exact (the three operations are order-defined point expressions; the solve in between is the mgSolve path, run
in the oracle with the correctly rounded RHS mean the device forms -- see tests/helpers.check_natural_mean_field)."""
import numpy as np
import pytest

from oracle import mgoracle as mo
from oracle.mgoracle import Oracle
from saras_b200 import capi
from tests.helpers import make_solver, max_abs

pytestmark = pytest.mark.gpu

CASES = [
    dict(name="nnn_uniform", xInd=4, yInd=4, zInd=4, vcDepth=2),
    dict(name="ppn_tanhz", xInd=5, yInd=4, zInd=4, vcDepth=3, xPer=1, yPer=1, zGrid=2, betaZ=1.5),
    dict(name="pnp_aniso", xInd=4, yInd=5, zInd=3, vcDepth=2, xPer=1, zPer=1, Lx=2.0, Ly=1.0, Lz=1.5),
    dict(name="tanh_all", xInd=4, yInd=4, zInd=5, vcDepth=3, xGrid=2, yGrid=2, zGrid=2, betaX=1.1, betaY=1.3, betaZ=1.7, Lx=2.0, Lz=1.5),
    dict(name="gs_nnp_65", xInd=6, yInd=6, zInd=6, vcDepth=4, zPer=1, gsSmooth=1),
]


def _fields(n, seed):
    rng = np.random.default_rng(seed)
    sh = mo.proj_shapes(n)
    return [np.ascontiguousarray(rng.uniform(-1.0, 1.0, sh[k])) for k in ("vx", "vy", "vz", "sc", "sc")]


def _setup(case, **kw):
    cfg = dict(dim=3, vcCount=4, mgTolerance=1e-9)
    cfg.update({k: v for k, v in case.items() if k != "name"})
    cfg.update(kw)
    o = Oracle(**cfg)
    s = make_solver(o)
    stag, coll, spacing = mo.proj_metrics(o)
    s.proj_setup(stag, coll, spacing)
    n, _ = o.level_dims(0)
    return o, s, n, stag, coll, spacing


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("general", [0, 1], ids=["pow2", "general_div"])
def test_divergence_bitexact(case, general, monkeypatch):
    if general:
        monkeypatch.setenv("SARASMG_NO_POW2", "1")
    o, s, n, stag, coll, spacing = _setup(case)
    vx, vy, vz, _, _ = _fields(n, 3)
    for which, a in ((capi.PROJ_VX, vx), (capi.PROJ_VY, vy), (capi.PROJ_VZ, vz)):
        s.proj_upload(which, a)
        assert np.array_equal(s.proj_download(which), a)
    for dt in (0.01, 1.0 / 3.0):
        s.proj_divergence(1.0 / dt)
        ref = mo.proj_divergence(n, vx, vy, vz, stag, spacing, 1.0 / dt)
        got = s.proj_download(capi.PROJ_RHS)
        assert np.array_equal(got, ref), max_abs(got - ref)
    s.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("general", [0, 1], ids=["pow2", "general_div"])
def test_correction_bitexact(case, general, monkeypatch):
    """Pp.syncData(); P += Pp; V -= dt * grad(Pp) with Pp = the level-0 pressure array the solver holds"""
    if general:
        monkeypatch.setenv("SARASMG_NO_POW2", "1")
    o, s, n, stag, coll, spacing = _setup(case)
    vx, vy, vz, P, pp = _fields(n, 5)
    for which, a in ((capi.PROJ_VX, vx), (capi.PROJ_VY, vy), (capi.PROJ_VZ, vz), (capi.PROJ_P, P)):
        s.proj_upload(which, a)
    s.level_set(0, capi.PRESSURE, pp)
    dt = 0.0125
    s.proj_correct(dt)
    mo.proj_correct(n, o.cfg["xPer"], o.cfg["yPer"], pp, P, vx, vy, vz, coll, spacing, dt)
    for which, ref, what in ((capi.PROJ_PP, pp, "Pp"), (capi.PROJ_P, P, "P"), (capi.PROJ_VX, vx, "Vx"), (capi.PROJ_VY, vy, "Vy"), (capi.PROJ_VZ, vz, "Vz")):
        got = s.proj_download(which)
        assert np.array_equal(got, ref), (what, max_abs(got - ref))
    s.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_projection_steps_bitexact(case):
    """three consecutive time steps' worth of divergence -> mgSolve -> correction with everything resident: Pp carries over as
    the next initial guess exactly as hydro_d3's member Pp does (hydro_d3.cc:343)"""
    o, s, n, stag, coll, spacing = _setup(case)
    o.set_exact_mean(True)
    vx, vy, vz, P, pp = _fields(n, 11)
    pp[...] = 0.0
    for which, a in ((capi.PROJ_VX, vx), (capi.PROJ_VY, vy), (capi.PROJ_VZ, vz), (capi.PROJ_P, P), (capi.PROJ_PP, pp)):
        s.proj_upload(which, a)
    dt = 0.01
    for step in range(3):
        got = s.proj_step(dt)
        rhs = mo.proj_divergence(n, vx, vy, vz, stag, spacing, 1.0 / dt)
        ref = o.mg_solve([pp], [rhs])
        mo.proj_correct(n, o.cfg["xPer"], o.cfg["yPer"], pp, P, vx, vy, vz, coll, spacing, dt)
        assert got["cycles"] == ref["cycles"], (step, got["cycles"], ref["cycles"])
        assert np.array_equal(got["coarse_iters"], ref["coarse_iters"])
        np.testing.assert_allclose(got["residuals"], ref["residuals"], rtol=1e-10, atol=0)
        for which, r, what in ((capi.PROJ_PP, pp, "Pp"), (capi.PROJ_P, P, "P"), (capi.PROJ_VX, vx, "Vx"), (capi.PROJ_VY, vy, "Vy"), (capi.PROJ_VZ, vz, "Vz")):
            g = s.proj_download(which)
            assert np.array_equal(g, r), (step, what, max_abs(g - r))
        # perturb the velocity as the next predictor step would, identically on both sides
        rng = np.random.default_rng(100 + step)
        for which, a in ((capi.PROJ_VX, vx), (capi.PROJ_VY, vy), (capi.PROJ_VZ, vz)):
            a += 0.1 * rng.uniform(-1.0, 1.0, a.shape)
            s.proj_upload(which, a)
    s.close()


def test_projection_reduces_divergence():
    """size-independent property: div(V_new) = div(V) - dt div(grad Pp) = dt * (rhs - Lap Pp) at every point whose stencil
    stays inside the core, i.e. the multigrid residual: after a converged solve the interior divergence collapses."""
    o, s, n, stag, coll, spacing = _setup(dict(name="nnn", xInd=5, yInd=5, zInd=5, vcDepth=3, gsSmooth=1), vcCount=12, mgTolerance=1e-12)
    vx, vy, vz, P, pp = _fields(n, 17)
    for which, a in ((capi.PROJ_VX, vx), (capi.PROJ_VY, vy), (capi.PROJ_VZ, vz)):
        s.proj_upload(which, a)
    s.proj_divergence(1.0)
    before = max_abs(s.proj_download(capi.PROJ_RHS)[2:-2, 2:-2, 2:-2])
    s.proj_step(0.01)
    s.proj_divergence(1.0)
    after = max_abs(s.proj_download(capi.PROJ_RHS)[2:-2, 2:-2, 2:-2])
    assert after < 1e-3 * before, (before, after)
    s.close()
