"""north_star: "the hydro_d2/d3 and scalar solvers in src/solvers call it unchanged".  oracle/_ref/libsarastsdropin_d3.so is the
reference's UNMODIFIED time stepper -- src/solvers/hydro/{hydro,hydro_d3}.cc, src/solvers/scalar/{scalar,scalar_d3}.cc, lib/field,
lib/boundary, lib/force, lib/initial, lib/grid, lib/parallel, lib/io/parser -- compiled and linked against THIS repository's
drop-in poisson class (saras_b200/host/poisson.{h,cc} over libsarasmg.so) in place of the reference's lib/poisson
(oracle/refbuild/Makefile).  CPU: it builds, loads, and refuses loudly without a GPU.  GPU: whole hydro_d3::timeAdvance() /
scalar_d3::timeAdvance() calls whose pressure solve runs on the device, against the same calls of the pure reference build."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import sarasts
from tests.golden import make_golden_timestep as gt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_libs = pytest.mark.skipif(not (sarasts.available() and sarasts.available(dropin=True)),
                                reason="oracle/_ref/libsarasts*_d3.so not built (needs /root/reference)")


def _has_gpu():
    try:
        import ctypes as C
        from saras_b200 import capi
        n = C.c_int(0)
        return capi.load().sarasmg_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@needs_libs
def test_time_stepper_links_against_the_dropin_and_exposes_the_same_entry_points():
    import ctypes as C
    lib = C.CDLL(sarasts.DROPIN_LIB)
    assert hasattr(lib, "sarasts_run") and hasattr(lib, "sarasts_extents")
    out = subprocess.run(["ldd", sarasts.DROPIN_LIB], capture_output=True, text=True).stdout
    assert "libsarasmg.so" in out        # the CUDA library is what its multigrid_d3 runs on


@needs_libs
def test_time_step_through_the_dropin_refuses_without_a_gpu():
    if _has_gpu():
        pytest.skip("a GPU is present")
    code = ("import numpy as np; from tests.golden import make_golden_timestep as gt; from oracle import sarasts; from oracle.mgoracle import Oracle;"
            "o = Oracle(**gt.CFG); p = sarasts.params(o.cfg, n_steps=1, output_only=gt.FIELDS, **gt.RUN); ext = sarasts.extents(p, True);"
            "sarasts.run(p, sarasts.TIME_ADVANCE, [{k: np.zeros(ext[k]) for k in gt.FIELDS}], dropin=True)")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "CUDA" in r.stderr        # no CPU fallback: message + abort, like the reference's fatal paths


@pytest.mark.gpu
@needs_libs
@pytest.mark.parametrize("case", sorted(gt.CASES))
def test_reference_time_steps_with_the_device_pressure_solve(case):
    """two timeAdvance() calls of the Taylor-Green vortex (hydro_d3) / of Rayleigh-Benard convection on a tanh mesh (scalar_d3):
    the reference build against the build whose mgSolver is this repository's.  The device forms the correctly rounded RHS mean,
    the reference a sequential sum, so the fields agree to the amplified rounding of that mean, not bit for
    bit: 1e-9 of the field's magnitude is asserted, ~1e-12 observed in the multigrid tests."""
    from oracle.mgoracle import Oracle
    cfg, run = gt.CASES[case]
    o = Oracle(**cfg)
    res = []
    for dropin in (False, True):
        p = sarasts.params(o.cfg, n_steps=2, output_only=gt.FIELDS, **run)
        ext = sarasts.extents(p, dropin)
        f = [{k: np.zeros(ext[k]) for k in gt.FIELDS}]
        if run.get("scalar"):
            f[0]["T"] = gt.rbc_temperature(ext["T"])
        sarasts.run(p, sarasts.TIME_ADVANCE, f, dropin=dropin)
        res.append(f[0])
    for k in ("vx", "vy", "vz", "mgRHS"):
        scale = max(float(np.abs(res[0][k]).max()), 1e-30)
        assert float(np.abs(res[0][k] - res[1][k]).max()) <= 1e-9 * scale, (k, float(np.abs(res[0][k] - res[1][k]).max()), scale)
    # P and Pp carry the arbitrary constant of the singular solve: compare them up to that constant
    for k in ("P", "Pp"):
        d = (res[0][k] - res[1][k])[1:-1, 1:-1, 1:-1]
        scale = max(float(np.abs(res[0][k]).max()), 1e-30)
        assert float(np.abs(d - d.mean()).max()) <= 1e-9 * scale, (k, float(np.abs(d - d.mean()).max()), scale)
