"""CPU checks of the implicit-solve oracle (oracle/impl_oracle.c).

* the squared metrics xix2Colloc .. ztz2Staggr it feeds the solves are pinned BIT FOR BIT against the reference's own grid
  object (oracle/_ref, grid.cc compiled unmodified);
* the Jacobi solve is checked through what it must deliver whatever the iteration details: the converged field satisfies
  (1 - dt/(2 Re) Lap) V = rhs in the bulk to cnTolerance (independent numpy statement of hydro_d3.cc:396-402), and the
  boundary objects leave the values lib/boundary/*.cc prescribe.  The pin against an EXECUTION of the reference's own
  hydro_d3::solveVx / Vy / Vz is tests/test_oracle_vs_timestepper.py."""
import os

import numpy as np
import pytest

from oracle import mgoracle as mo, sarasref
from oracle.mgoracle import Oracle
from saras_b200 import capi

HAVE_REF = os.path.isdir(sarasref.REFERENCE_SRC) or sarasref.available(3, 0)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("g", [dict(xInd=4, yInd=4, zInd=4, vcDepth=2),
                               dict(xInd=5, yInd=4, zInd=3, vcDepth=2, xGrid=2, yGrid=2, zGrid=2, betaX=1.1, betaY=1.3, betaZ=1.7, Lx=2.0, Ly=1.0, Lz=1.5)],
                         ids=["uniform", "tanh"])
def test_squared_metrics_match_reference_grid(g):
    ref = sarasref.proj_metrics(dim=3, **g)[0]
    coll2, stag2 = mo.impl_metrics2(Oracle(dim=3, **g))
    for d, (cn, sn) in enumerate((("xix2Colloc", "xix2Staggr"), ("ety2Colloc", "ety2Staggr"), ("ztz2Colloc", "ztz2Staggr"))):
        assert np.array_equal(coll2[d], ref[cn]), cn
        assert np.array_equal(stag2[d], ref[sn]), sn


def _bulk(n, comp, per):
    sl = []
    for d in range(3):
        stag = d != comp
        cs = n[d] if stag else n[d] - 1
        cut = 1 if (stag and not per[d]) else 0
        sl.append((cut, cs - 1 - cut))
    return sl


@pytest.mark.parametrize("case", [dict(name="ldc", per=(0, 0, 0), prob=1), dict(name="ppp", per=(1, 1, 1), prob=5),
                                  dict(name="channel_ppn_tanh", per=(1, 1, 0), prob=4, zGrid=2, betaZ=1.4),
                                  dict(name="inflow", per=(0, 0, 0), prob=3)], ids=lambda c: c["name"])
def test_converged_solve_satisfies_the_implicit_equation(case):
    extra = {k: v for k, v in case.items() if k not in ("name", "per", "prob")}
    per = case["per"]
    o = Oracle(dim=3, xInd=4, yInd=3, zInd=4, vcDepth=1, xPer=per[0], yPer=per[1], zPer=per[2], **extra)
    t, v = capi.velocity_bcs(per[0], per[1], per[2], case["prob"])
    dt, Re, tol = 0.02, 2.0, 1e-9
    io = mo.ImplOracle(o, dt, Re, tol, t, v)
    n = io.n
    sh = mo.proj_shapes(n)
    rng = np.random.default_rng(3)
    hx, hy, hz = io.spacing
    for comp, key in enumerate(("vx", "vy", "vz")):
        V = np.ascontiguousarray(rng.uniform(-1, 1, sh[key]))
        rhs = np.ascontiguousarray(rng.uniform(-1, 1, sh[key]))
        sweeps, err = io.solve(comp, V, rhs)
        assert sweeps > 3 and err < tol
        (i0, i1), (j0, j1), (k0, k1) = _bulk(n, comp, per)
        c = V[i0 + 1:i1 + 2, j0 + 1:j1 + 2, k0 + 1:k1 + 2]
        m = [(io.stag2[d] if d != comp else io.coll2[d])[1:] for d in range(3)]          # [i] = index i
        lap = (m[0][i0:i1 + 1, None, None] * (V[i0 + 2:i1 + 3, j0 + 1:j1 + 2, k0 + 1:k1 + 2] - 2.0 * c + V[i0:i1 + 1, j0 + 1:j1 + 2, k0 + 1:k1 + 2]) / (hx * hx) +
               m[1][None, j0:j1 + 1, None] * (V[i0 + 1:i1 + 2, j0 + 2:j1 + 3, k0 + 1:k1 + 2] - 2.0 * c + V[i0 + 1:i1 + 2, j0:j1 + 1, k0 + 1:k1 + 2]) / (hy * hy) +
               m[2][None, None, k0:k1 + 1] * (V[i0 + 1:i1 + 2, j0 + 1:j1 + 2, k0 + 2:k1 + 3] - 2.0 * c + V[i0 + 1:i1 + 2, j0 + 1:j1 + 2, k0:k1 + 1]) / (hz * hz))
        resid = np.abs(c - lap * 0.5 * dt * (1.0 / Re) - rhs[i0 + 1:i1 + 2, j0 + 1:j1 + 2, k0 + 1:k1 + 2])
        assert float(resid.max()) == err                     # the solve's own last maxError, reproduced independently
        # boundary objects (lib/boundary/*.cc:71-84)
        for w in range(6):
            ax, hi = w // 2, w & 1
            lo_b, hi_b = _bulk(n, comp, per)[ax]
            wall = (hi_b + 1 if hi else lo_b - 1) + 1           # array index of the wall plane
            take = lambda idx: np.take(V, idx, axis=ax)
            # the later walls overwrite the rims of the earlier ones: check the part no later wall touches
            core = tuple(slice(2, -2) for _ in range(2))
            inner = wall + (-1 if hi else 1)
            if t[comp][w] == capi.BC_DIRICHLET_CC:
                assert np.all(take(wall)[core] == v[comp][w])
            elif t[comp][w] == capi.BC_DIRICHLET_FC:
                assert np.array_equal(take(wall)[core], (2.0 * v[comp][w] - take(inner))[core])
            elif t[comp][w] == capi.BC_NEUMANN:
                assert np.array_equal(take(wall)[core], take(inner)[core])
