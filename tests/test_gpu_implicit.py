"""Implicit velocity solves on the device: hydro_d3::solveVx / solveVy / solveVz with imposeUBCs / VBCs / WBCs
through the C-ABI (sarasmg_impl_*) against oracle/impl_oracle.c on identical seeded fields.  Bar: bit-exact fields, identical
sweep counts, identical final max-norm (a max is order-independent)."""
import numpy as np
import pytest

from oracle import mgoracle as mo
from oracle.mgoracle import Oracle
from saras_b200 import capi
from tests.helpers import make_solver, max_abs

pytestmark = pytest.mark.gpu

CASES = [
    dict(name="ldc_nnn", per=(0, 0, 0), prob=1, xInd=4, yInd=4, zInd=4),
    dict(name="ppp_aniso", per=(1, 1, 1), prob=5, xInd=4, yInd=3, zInd=5, Lx=2.0, Lz=1.5),
    dict(name="channel_ppn_tanhz", per=(1, 1, 0), prob=4, xInd=4, yInd=4, zInd=4, zGrid=2, betaZ=1.4),
    dict(name="inflow_tanh_all", per=(0, 0, 0), prob=3, xInd=4, yInd=4, zInd=5, xGrid=2, yGrid=2, zGrid=2, betaX=1.1, betaY=1.3, betaZ=1.7),
]


def _setup(case, dt=0.02, Re=2.0, tol=1e-9):
    per = case["per"]
    extra = {k: v for k, v in case.items() if k not in ("name", "per", "prob")}
    o = Oracle(dim=3, vcDepth=2, vcCount=3, xPer=per[0], yPer=per[1], zPer=per[2], **extra)
    s = make_solver(o)
    stag, coll, spacing = mo.proj_metrics(o)
    s.proj_setup(stag, coll, spacing)
    t, v = capi.velocity_bcs(per[0], per[1], per[2], case["prob"])
    io = mo.ImplOracle(o, dt, Re, tol, t, v)
    s.impl_setup(dt, Re, tol, spacing, io.coll2, io.stag2, t, v, io.max_iterations)
    return o, s, io


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_velocity_bcs_bitexact(case):
    o, s, io = _setup(case)
    sh = mo.proj_shapes(io.n)
    rng = np.random.default_rng(21)
    for comp, key in enumerate(("vx", "vy", "vz")):
        V = np.ascontiguousarray(rng.uniform(-1, 1, sh[key]))
        s.proj_upload(comp, V)
        s.impl_bcs(comp)
        io.impose_bcs(comp, V)
        got = s.proj_download(comp)
        assert np.array_equal(got, V), (comp, max_abs(got - V))
    s.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("general", [0, 1], ids=["pow2", "general_div"])
def test_implicit_solves_bitexact(case, general, monkeypatch):
    if general:         # the divisors hx*hx .. are powers of two on every SARAS grid; force the exact-reciprocal sequence instead
        monkeypatch.setenv("SARASMG_NO_POW2", "1")
    o, s, io = _setup(case)
    sh = mo.proj_shapes(io.n)
    rng = np.random.default_rng(22)
    for comp, key in enumerate(("vx", "vy", "vz")):
        V = np.ascontiguousarray(rng.uniform(-1, 1, sh[key]))
        rhs = np.ascontiguousarray(rng.uniform(-1, 1, sh[key]))
        s.proj_upload(comp, V)
        s.impl_upload_rhs(comp, rhs)
        sweeps, err = s.impl_solve(comp)
        ref_sweeps, ref_err = io.solve(comp, V, rhs)
        assert ref_sweeps > 3
        assert sweeps == ref_sweeps, (comp, sweeps, ref_sweeps)
        assert err == ref_err, (comp, err, ref_err)
        got = s.proj_download(comp)
        assert np.array_equal(got, V), (comp, max_abs(got - V))
    s.close()


def test_predictor_then_projection_on_the_device():
    """solveVx/Vy/Vz followed by the projection step, everything resident: the tail of hydro_d3::timeAdvance (hydro_d3.cc:306-362)
    from nseRHS on, against the three oracles chained the same way"""
    case = CASES[0]
    o, s, io = _setup(case, Re=5.0)
    o.set_exact_mean(True)
    n = io.n
    sh = mo.proj_shapes(n)
    stag, coll, spacing = mo.proj_metrics(o)
    rng = np.random.default_rng(23)
    V = [np.ascontiguousarray(rng.uniform(-1, 1, sh[k])) for k in ("vx", "vy", "vz")]
    R = [np.ascontiguousarray(rng.uniform(-1, 1, sh[k])) for k in ("vx", "vy", "vz")]
    P = np.ascontiguousarray(rng.uniform(-1, 1, sh["sc"]))
    pp = np.zeros(sh["sc"])
    for comp in range(3):
        s.proj_upload(comp, V[comp])
        s.impl_upload_rhs(comp, R[comp])
    s.proj_upload(capi.PROJ_P, P)
    s.proj_upload(capi.PROJ_PP, pp)
    dt = 0.02
    for comp in range(3):
        a, _ = s.impl_solve(comp)
        b, _ = io.solve(comp, V[comp], R[comp])
        assert a == b
    got = s.proj_step(dt)
    rhs = mo.proj_divergence(n, V[0], V[1], V[2], stag, spacing, 1.0 / dt)
    ref = o.mg_solve([pp], [rhs])
    mo.proj_correct(n, 0, 0, pp, P, V[0], V[1], V[2], coll, spacing, dt)
    assert got["cycles"] == ref["cycles"]
    for which, r in ((capi.PROJ_VX, V[0]), (capi.PROJ_VY, V[1]), (capi.PROJ_VZ, V[2]), (capi.PROJ_P, P), (capi.PROJ_PP, pp)):
        g = s.proj_download(which)
        assert np.array_equal(g, r), (which, max_abs(g - r))
    s.close()
