"""The reference's I/O-less full time step in the oracle build (oracle side) and the fixture captured from it.

oracle/_ref/libsarasts_d3.so is the reference's unmodified hydro_d3 time stepper (oracle/refbuild/ts_driver.cc).  Its computeDiff /
computeNLin go through the stand-in Blitz stencil operators (central12n / central22n), which nothing else pins; they are checked
here against what they must compute: second-order convergence to the analytic Laplacian, a divergence-free Taylor-Green vortex
that decays at the viscous rate and stays divergence-free after the projection.  Then tests/golden/tgv_step_33.npz -- two real
time steps of the Taylor-Green vortex -- pins the multigrid oracle on right-hand sides the time stepper itself produced."""
import os

import numpy as np
import pytest

from oracle import mgoracle as mo, sarasts
from oracle.mgoracle import Oracle
from tests.golden import make_golden_timestep as gt

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ts = pytest.mark.skipif(not sarasts.available(), reason="oracle/_ref/libsarasts_d3.so not built (needs /root/reference)")


@needs_ts
def test_reference_compute_diff_converges_to_the_analytic_laplacian():
    """vfield::computeDiff (Crank-Nicolson scheme: HALF the Laplacian, derivative.cc:185-186) of sin(kx) cos(ky) cos(2kz) on a
    periodic unit box: relative error falls by 4 per refinement -- an index misalignment in the stencils would be O(1)"""
    errs = []
    k = 2.0 * np.pi
    for ind in (4, 5, 6):
        o = Oracle(dim=3, xInd=ind, yInd=ind, zInd=ind, vcDepth=2, xPer=1, yPer=1, zPer=1)
        p = sarasts.params(o.cfg, prob_type=5, dt=0.001, Re=100.0)
        ext = sarasts.extents(p)
        h = 1.0 / (1 << ind)

        def coords(shape, comp):
            c = [((np.arange(-1, shape[d] - 1) + 0.5) * h if d == comp else np.arange(-1, shape[d] - 1) * h) for d in range(3)]
            return np.meshgrid(*c, indexing="ij")

        flds = {}
        for comp, name in enumerate(("vx", "vy", "vz")):
            X, Y, Z = coords(ext[name], comp)
            flds[name] = np.ascontiguousarray(np.sin(k * X) * np.cos(k * Y) * np.cos(2 * k * Z))
        for nm in ("rhs_x", "rhs_y", "rhs_z"):
            flds[nm] = np.zeros(ext[nm])
        sarasts.run(p, sarasts.COMPUTE_DIFF, [flds])
        for comp, (name, out) in enumerate((("vx", "rhs_x"), ("vy", "rhs_y"), ("vz", "rhs_z"))):
            X, Y, Z = coords(ext[name], comp)
            exact = -0.5 * 6.0 * k * k * np.sin(k * X) * np.cos(k * Y) * np.cos(2 * k * Z)
            core = (slice(1, -1),) * 3
            errs.append((ind, comp, float(np.abs(flds[out][core] - exact[core]).max() / np.abs(exact).max())))
    for comp in range(3):
        e = [x[2] for x in errs if x[1] == comp]
        assert e[0] < 0.05 and 3.5 < e[0] / e[1] < 4.5 and 3.5 < e[1] / e[2] < 4.5, errs


@needs_ts
def test_reference_convective_term_converges_to_the_analytic_one():
    """vfield::computeNLin of the Taylor-Green velocity field, x component: -(u u_x + v u_y + w u_z) to second order.  This is
    the check of the NORMALISED first-derivative stencil (central12n = (A(+1) - A(-1)) / 2): without the 1/2 the ratio to the
    analytic term would tend to 2, not 1."""
    L = 2.0 * np.pi
    errs = []
    for ind in (4, 5, 6):
        o = Oracle(dim=3, xInd=ind, yInd=ind, zInd=ind, vcDepth=2, xPer=1, yPer=1, zPer=1, Lx=L, Ly=L, Lz=L)
        p = sarasts.params(o.cfg, prob_type=5, dt=0.001, Re=100.0)
        ext = sarasts.extents(p)
        h = L / (1 << ind)

        def coords(shape, comp):
            c = [((np.arange(-1, shape[d] - 1) + 0.5) * h if d == comp else np.arange(-1, shape[d] - 1) * h) for d in range(3)]
            return np.meshgrid(*c, indexing="ij")

        f = {}
        X, Y, Z = coords(ext["vx"], 0)
        f["vx"] = np.ascontiguousarray(np.sin(X) * np.cos(Y) * np.cos(Z))
        X, Y, Z = coords(ext["vy"], 1)
        f["vy"] = np.ascontiguousarray(-np.cos(X) * np.sin(Y) * np.cos(Z))
        f["vz"] = np.zeros(ext["vz"])
        for nm in ("rhs_x", "rhs_y", "rhs_z"):
            f[nm] = np.zeros(ext[nm])
        sarasts.run(p, sarasts.COMPUTE_NLIN, [f])
        X, Y, Z = coords(ext["vx"], 0)
        conv = (np.sin(X) * np.cos(Y) * np.cos(Z)) * (np.cos(X) * np.cos(Y) * np.cos(Z)) + \
               (-np.cos(X) * np.sin(Y) * np.cos(Z)) * (-np.sin(X) * np.sin(Y) * np.cos(Z))
        core = (slice(2, -2),) * 3
        errs.append(float(np.abs(f["rhs_x"][core] + conv[core]).max() / np.abs(conv).max()))
    assert errs[0] < 0.08 and 3.5 < errs[0] / errs[1] < 4.5 and 3.5 < errs[1] / errs[2] < 4.5, errs


@needs_ts
def test_reference_taylor_green_steps_are_physical():
    """hydro_d3::timeAdvance x 3 from the reference's own Taylor-Green initial condition (Re = 100, 2 pi box): divergence-free
    start, kinetic energy decaying as exp(-2 nu k^2 t) with k^2 = 3 (the nonlinear term only redistributes energy at this
    early time), projected velocity divergence-free to the multigrid tolerance"""
    L = 2.0 * np.pi
    o = Oracle(dim=3, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=10, xPer=1, yPer=1, zPer=1, gsSmooth=1, mgTolerance=1e-6, Lx=L, Ly=L, Lz=L)
    n, _ = o.level_dims(0)
    stag, _, sp = mo.proj_metrics(o)
    nu, dt = 0.01, 0.01

    def state(nsteps):
        p = sarasts.params(o.cfg, prob_type=5, dt=dt, Re=1.0 / nu, cn_tolerance=1e-6, ic_type=1, n_steps=nsteps, output_only=gt.FIELDS)
        ext = sarasts.extents(p)
        f = [{k: np.zeros(ext[k]) for k in gt.FIELDS}]
        sarasts.run(p, sarasts.TIME_ADVANCE, f)
        return f[0]

    def energy(x):
        c = (slice(1, -2),) * 3            # one period: the last staggered point repeats the first
        return 0.5 * sum(float(np.mean(x[k][1:-1, 1:-1, 1:-1] ** 2)) for k in ("vx", "vy", "vz"))

    def max_div(x):
        return float(np.abs(mo.proj_divergence(n, x["vx"], x["vy"], x["vz"], stag, sp, 1.0)[2:-2, 2:-2, 2:-2]).max())

    s0, s3 = state(0), state(3)
    assert max_div(s0) < 1e-13
    assert max_div(s3) < 1e-3
    decay = energy(s3) / energy(s0)
    assert abs(decay - np.exp(-2.0 * nu * 3.0 * 3 * dt)) < 2e-4, decay


@pytest.mark.parametrize("case", sorted(gt.CASES))
def test_multigrid_oracle_reproduces_the_captured_time_steps(case):
    """tests/golden/{tgv,rbc}_step_33.npz (two timeAdvance() calls of the reference each, tests/golden/make_golden_timestep.py): from
    the right-hand side the time stepper formed and the previous Pp as initial guess, the multigrid oracle + Pp.syncData() must
    give the Pp the reference was left with, bit for bit"""
    g = np.load(os.path.join(HERE, "golden", case + ".npz"))
    o = Oracle(**gt.CASES[case][0])
    pp = np.zeros_like(g["Pp_1"])
    for step in (1, 2):
        o.mg_solve([pp], [np.ascontiguousarray(g["mgRHS_%d" % step])])
        mo.proj_sync_multi(o, [pp])
        assert np.array_equal(pp, g["Pp_%d" % step]), (step, float(np.abs(pp - g["Pp_%d" % step]).max()))


@needs_ts
@pytest.mark.parametrize("case", sorted(gt.CASES))
def test_time_step_fixture_is_what_the_reference_produces_now(case):
    g = np.load(os.path.join(HERE, "golden", case + ".npz"))
    r = gt.reference_steps(2, case)
    assert np.array_equal(r["mgRHS"], g["mgRHS_2"]) and np.array_equal(r["Pp"], g["Pp_2"])
