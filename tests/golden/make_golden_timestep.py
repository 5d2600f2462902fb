"""Generates tests/golden/tgv_step_33.npz and tests/golden/rbc_step_33.npz from the REFERENCE's own time stepper (oracle/_ref/libsarasts_d3.so: the unmodified
hydro_d3 / lib/field / lib/boundary / lib/poisson sources, oracle/refbuild/ts_driver.cc): two full hydro_d3::timeAdvance() calls
of the Taylor-Green vortex (BASELINE config C3's problem: Re = 1600, periodic 2 pi box) on a 33^3 grid, Gauss-Seidel multigrid as
in the reference's YAML.  Stored per step: mgRHS = div(V*) / dt as the time stepper formed it (after its own computeDiff /
computeNLin / implicit solves) and Pp after mgSolve + syncData -- real right-hand sides of the pressure solve instead of
surrogates.  rbc_step_33.npz: the same from scalar_d3::timeAdvance() for BASELINE config C4's problem -- Rayleigh-Benard convection,
Ra = 1e8, Pr = 1, free-fall units, buoyancy forcing, tanh-stretched z between the plates, periodic x / y -- started from the
conduction profile plus a seeded temperature perturbation (stored in the fixture).

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


RBC_CFG = dict(dim=3, xInd=5, yInd=5, zInd=5, vcDepth=3, vcCount=10, xPer=1, yPer=1, zPer=0, gsSmooth=1, resType=2, mgTolerance=1.0e-6,
               zGrid=2, betaZ=1.3)
RBC_RUN = dict(prob_type=5, dt=0.01, Re=1000.0, cn_tolerance=1.0e-5, scalar=True, rbc_type=2, force_type=3, Ra=1.0e8, Pr=1.0)
CASES = {"tgv_step_33": (CFG, RUN), "rbc_step_33": (RBC_CFG, RBC_RUN)}


def rbc_temperature(shape):
    rng = np.random.default_rng(3)
    return np.ascontiguousarray(0.5 + 0.1 * rng.uniform(-1.0, 1.0, shape))


def reference_steps(nsteps, case="tgv_step_33"):
    cfg, run = CASES[case]
    o = Oracle(**cfg)
    p = sarasts.params(o.cfg, n_steps=nsteps, output_only=FIELDS, **run)
    ext = sarasts.extents(p)
    f = [{k: np.zeros(ext[k]) for k in FIELDS}]
    if run.get("scalar"):
        f[0]["T"] = rbc_temperature(ext["T"])
    sarasts.run(p, sarasts.TIME_ADVANCE, f)
    return f[0]


def main():
    for case in CASES:
        out = {}
        for step in (1, 2):
            r = reference_steps(step, case)
            out["mgRHS_%d" % step] = r["mgRHS"]
            out["Pp_%d" % step] = r["Pp"]
            print("%s step %d: max |mgRHS| = %.6e, max |Pp| = %.6e" % (case, step, np.abs(r["mgRHS"]).max(), np.abs(r["Pp"]).max()))
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **out)


if __name__ == "__main__":
    main()
