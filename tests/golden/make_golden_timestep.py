"""Generates tests/golden/tgv_step_33.npz from the REFERENCE's own time stepper (oracle/_ref/libsarasts_d3.so: the unmodified
hydro_d3 / lib/field / lib/boundary / lib/poisson sources, oracle/refbuild/ts_driver.cc): two full hydro_d3::timeAdvance() calls
of the Taylor-Green vortex (BASELINE config C3's problem: Re = 1600, periodic 2 pi box) on a 33^3 grid, Gauss-Seidel multigrid as
in the reference's YAML.  Stored per step: mgRHS = div(V*) / dt as the time stepper formed it (after its own computeDiff /
computeNLin / implicit solves) and Pp after mgSolve + syncData -- real right-hand sides of the pressure solve instead of
surrogates.

    python -m tests.golden.make_golden_timestep        (only where /root/reference exists; the fixture is committed)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import sarasts  # noqa: E402
from oracle.mgoracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
L = 2.0 * np.pi
CFG = dict(dim=3, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=10, xPer=1, yPer=1, zPer=1, gsSmooth=1, resType=2, mgTolerance=1.0e-6,
           Lx=L, Ly=L, Lz=L)
RUN = dict(prob_type=5, dt=0.01, Re=1600.0, cn_tolerance=1.0e-5, ic_type=1)
FIELDS = ("vx", "vy", "vz", "P", "Pp", "mgRHS")


def reference_steps(nsteps):
    o = Oracle(**CFG)
    p = sarasts.params(o.cfg, n_steps=nsteps, output_only=FIELDS, **RUN)
    ext = sarasts.extents(p)
    f = [{k: np.zeros(ext[k]) for k in FIELDS}]
    sarasts.run(p, sarasts.TIME_ADVANCE, f)
    return f[0]


def main():
    out = {}
    for step in (1, 2):
        r = reference_steps(step)
        out["mgRHS_%d" % step] = r["mgRHS"]
        out["Pp_%d" % step] = r["Pp"]
        print("step %d: max |mgRHS| = %.6e, max |Pp| = %.6e" % (step, np.abs(r["mgRHS"]).max(), np.abs(r["Pp"]).max()))
    np.savez_compressed(os.path.join(HERE, "tgv_step_33.npz"), **out)


if __name__ == "__main__":
    main()
