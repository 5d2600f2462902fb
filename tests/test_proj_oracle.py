"""CPU checks of the projection-step oracle (oracle/proj_oracle.c).

* its restatement of the first-derivative metrics (grid.cc:370-519) is pinned BIT FOR BIT against the reference's own grid
  object (oracle/_ref, compiled unmodified) on uniform and tanh-stretched grids, 1 / 2 / 4 ranks;
* its divergence / syncData / correction loops are checked against an independent vectorised numpy statement of the Blitz
  expressions (vfield.cc:310-319, plainsf.h:92-98, sfield.cc:197-201, mpidata.cc:105-259).  The reference's lib/field cannot
  (vfield.cc:310-319, plainsf.h:92-98, sfield.cc:197-201, mpidata.cc:105-259) and the global periodic wrap.  The pin against an
  EXECUTION of the reference's own field library is tests/test_oracle_vs_timestepper.py."""
import os

import numpy as np
import pytest

from oracle import mgoracle as mo, sarasref
from oracle.mgoracle import Oracle

HAVE_REF = os.path.isdir(sarasref.REFERENCE_SRC) or sarasref.available(3, 0)

GRIDS = [
    dict(xInd=4, yInd=4, zInd=4, vcDepth=2),
    dict(xInd=5, yInd=4, zInd=3, vcDepth=2, xGrid=2, yGrid=2, zGrid=2, betaX=1.1, betaY=1.3, betaZ=1.7, Lx=2.0, Ly=1.0, Lz=1.5),
    dict(xInd=5, yInd=5, zInd=4, vcDepth=2, npX=2, npY=2, zGrid=2, betaZ=1.5, xGrid=2, betaX=0.9),
    dict(xInd=5, yInd=4, zInd=4, vcDepth=2, npX=2, npY=1, yGrid=2, betaY=2.0, Ly=3.0),
]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("g", GRIDS, ids=lambda g: "x%d_y%d_z%d_np%dx%d" % (g["xInd"], g["yInd"], g["zInd"], g.get("npX", 1), g.get("npY", 1)))
def test_first_derivative_metrics_match_reference_grid(g):
    ref = sarasref.proj_metrics(dim=3, **g)
    o = Oracle(dim=3, **g)
    for r in range(o.nranks):
        stag, coll, spacing = mo.proj_metrics(o, r)
        assert tuple(spacing) == tuple(ref[r]["spacing"])
        for d, (sn, cn) in enumerate((("xi_xStaggr", "xi_xColloc"), ("et_yStaggr", "et_yColloc"), ("zt_zStaggr", "zt_zColloc"))):
            assert np.array_equal(stag[d], ref[r][sn]), (r, sn, stag[d], ref[r][sn])
            assert np.array_equal(coll[d], ref[r][cn]), (r, cn, coll[d], ref[r][cn])


def _np_divergence(vx, vy, vz, stag, sp, inv_dt):
    xs, ys, zs = (a[1:-1] for a in stag)
    rhs = np.zeros((vx.shape[0] + 1, vx.shape[1], vx.shape[2]))
    c = (slice(1, -1),) * 3
    a = xs[:, None, None] * (vx[1:, 1:-1, 1:-1] - vx[:-1, 1:-1, 1:-1]) / sp[0]
    b = ys[None, :, None] * (vy[1:-1, 1:, 1:-1] - vy[1:-1, :-1, 1:-1]) / sp[1]
    d = zs[None, None, :] * (vz[1:-1, 1:-1, 1:] - vz[1:-1, 1:-1, :-1]) / sp[2]
    rhs[c] = (a + b) + d
    return rhs * inv_dt


def _np_correct(xper, yper, pp, P, vx, vy, vz, coll, sp, dt):
    pp = pp.copy()
    if xper:
        lo, hi = pp[-3, 1:-1, 1:-1].copy(), pp[2, 1:-1, 1:-1].copy()
        pp[0, 1:-1, 1:-1], pp[-1, 1:-1, 1:-1] = lo, hi
    if yper:
        lo, hi = pp[1:-1, -3, 1:-1].copy(), pp[1:-1, 2, 1:-1].copy()
        pp[1:-1, 0, 1:-1], pp[1:-1, -1, 1:-1] = lo, hi
    c = [pp[-3, -3, 1:-1].copy(), pp[-3, 2, 1:-1].copy(), pp[2, -3, 1:-1].copy(), pp[2, 2, 1:-1].copy()]
    pp[0, 0, 1:-1], pp[0, -1, 1:-1], pp[-1, 0, 1:-1], pp[-1, -1, 1:-1] = c
    P = P + pp
    xc, yc, zc = (a[1:-1] for a in coll)          # core of the collocated metric: indices 0 .. n-2
    core = pp[1:-1, 1:-1, 1:-1]
    vx, vy, vz = vx.copy(), vy.copy(), vz.copy()
    vx[1:-1, 1:-1, 1:-1] -= (xc[:, None, None] * (core[1:] - core[:-1]) / sp[0]) * dt
    vy[1:-1, 1:-1, 1:-1] -= (yc[None, :, None] * (core[:, 1:] - core[:, :-1]) / sp[1]) * dt
    vz[1:-1, 1:-1, 1:-1] -= (zc[None, None, :] * (core[:, :, 1:] - core[:, :, :-1]) / sp[2]) * dt
    return pp, P, vx, vy, vz


@pytest.mark.parametrize("g", GRIDS[:2] + [dict(xInd=4, yInd=3, zInd=5, vcDepth=2, xPer=1, yPer=1), dict(xInd=3, yInd=4, zInd=4, vcDepth=2, yPer=1, zGrid=2, betaZ=1.2)],
                         ids=["uniform", "tanh", "ppn", "npn_tanhz"])
def test_projection_oracle_against_numpy(g):
    o = Oracle(dim=3, **g)
    n, _ = o.level_dims(0)
    stag, coll, sp = mo.proj_metrics(o)
    rng = np.random.default_rng(9)
    sh = mo.proj_shapes(n)
    vx, vy, vz, P, pp = [np.ascontiguousarray(rng.uniform(-1, 1, sh[k])) for k in ("vx", "vy", "vz", "sc", "sc")]
    dt = 0.0125
    assert np.array_equal(mo.proj_divergence(n, vx, vy, vz, stag, sp, 1.0 / dt), _np_divergence(vx, vy, vz, stag, sp, 1.0 / dt))
    want = _np_correct(o.cfg["xPer"], o.cfg["yPer"], pp, P, vx, vy, vz, coll, sp, dt)
    mo.proj_correct(n, o.cfg["xPer"], o.cfg["yPer"], pp, P, vx, vy, vz, coll, sp, dt)
    for got, ref, what in zip((pp, P, vx, vy, vz), want, ("Pp", "P", "Vx", "Vy", "Vz")):
        assert np.array_equal(got, ref), what


@pytest.mark.parametrize("lay", [(2, 1, 0, 0), (2, 2, 1, 1), (4, 2, 1, 0), (1, 2, 0, 1)], ids=lambda t: "np%dx%d_per%d%d" % t)
def test_multirank_sync_is_the_global_periodic_wrap(lay):
    """mpidata::syncData across ranks (projo_sync_multi) against a direct statement on the GLOBAL staggered field: every rank's
    core is a window of one global array whose neighbouring windows share their boundary plane; after the exchange a pad holds
    the global value one point outside the window -- across rank boundaries, and across the domain boundary where the
    reference wraps (periodic faces; edges always, parallel.cc:124-129) -- and is untouched at wall faces."""
    npx, npy, xper, yper = lay
    o = Oracle(dim=3, xInd=5, yInd=5, zInd=3, vcDepth=1, npX=npx, npY=npy, xPer=xper, yPer=yper)
    n, F = o.level_dims(0)
    gx, gy = npx * (n[0] - 1), npy * (n[1] - 1)           # global intervals: global staggered points 0 .. gx, point gx == point 0 when periodic
    rng = np.random.default_rng(5)
    G = rng.uniform(-1, 1, (gx + 1, gy + 1, n[2]))
    if xper:
        G[gx] = G[0]
    if yper:
        G[:, gy] = G[:, 0]
    arrs, before = [], []
    for r in range(o.nranks):
        xr, yr = r % npx, r // npx
        a = rng.uniform(-1, 1, F)
        a[1:-1, 1:-1, 1:-1] = G[xr * (n[0] - 1):xr * (n[0] - 1) + n[0], yr * (n[1] - 1):yr * (n[1] - 1) + n[1], :]
        arrs.append(np.ascontiguousarray(a)); before.append(a.copy())
    mo.proj_sync_multi(o, arrs)

    def wrap(i, g):     # global staggered index of a point one step outside, on a ring of g intervals
        return i % g

    for r in range(o.nranks):
        xr, yr = r % npx, r // npx
        x0, y0 = xr * (n[0] - 1), yr * (n[1] - 1)
        a = arrs[r]
        assert np.array_equal(a[1:-1, 1:-1, :], before[r][1:-1, 1:-1, :])                   # cores and z pads untouched
        for side, (li, gi) in enumerate(((0, x0 - 1), (n[0] + 1, x0 + n[0]))):
            wall = not xper and ((side == 0 and xr == 0) or (side == 1 and xr == npx - 1))
            want = before[r][li, 1:-1, 1:-1] if wall else G[wrap(gi, gx), y0:y0 + n[1], :]
            assert np.array_equal(a[li, 1:-1, 1:-1], want), (r, "x", side)
        for side, (lj, gj) in enumerate(((0, y0 - 1), (n[1] + 1, y0 + n[1]))):
            wall = not yper and ((side == 0 and yr == 0) or (side == 1 and yr == npy - 1))
            want = before[r][1:-1, lj, 1:-1] if wall else G[x0:x0 + n[0], wrap(gj, gy), :]
            assert np.array_equal(a[1:-1, lj, 1:-1], want), (r, "y", side)
        if xper and yper:       # fully periodic: the edge pads are the global diagonal points too
            for li, gi in ((0, x0 - 1), (n[0] + 1, x0 + n[0])):
                for lj, gj in ((0, y0 - 1), (n[1] + 1, y0 + n[1])):
                    assert np.array_equal(a[li, lj, 1:-1], G[wrap(gi, gx), wrap(gj, gy), :]), (r, "edge")
