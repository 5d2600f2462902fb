"""GPU parity at the sizes BASELINE.json names.

C1 (129^3 mgTest), C2 (257^2 LDC shape) and C3 (257^3 PPP, TGV shape) are small enough for the CPU oracle to finish in
seconds, so they are compared with it directly: identical cycle and coarse-iteration counts, residual history, field.
C4 (513^3, tanh-stretched) is compared with the oracle directly too (2 cycles, about a minute of CPU per solve), and,
like C5 (1025^3: a CPU run would need ~40 GB and hours), through size-independent properties of the operator that hold
BIT FOR BIT in IEEE arithmetic:
  * scaling: every operation of the V-cycle commutes with multiplication by 2 (no rounding is involved), so with the
    data-dependent stopping tests disabled mgSolve(2 rhs) == 2 mgSolve(rhs) exactly;
  * two kernel families: the TMA-staged and the register-marching stencil kernels are independent implementations and
    must produce identical bits at full size;
  * repeatability: the persistent device hierarchy carries no state between calls.
"""
import numpy as np
import pytest

from oracle.mgoracle import Oracle
from saras_b200 import host_api
from tests.helpers import make_solver, seeded_field, smooth_rhs

pytestmark = pytest.mark.gpu


def _host_mem_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 1e9
    except Exception:
        return 0.0


def test_c1_mgtest_129_against_oracle():
    """tests/mgTest/input/parameters.yaml: 129^3, depth 5, 25 V(2,2) cycles, Gauss-Seidel, resType 2, TEST_POISSON (rhs = 1,
    Dirichlet ghost r^2/6); BASELINE config 0 runs it on one rank."""
    cfg = dict(dim=3, test_poisson=1, xInd=7, yInd=7, zInd=7, vcDepth=5, vcCount=25, gsSmooth=1, resType=2, mgTolerance=1e-6)
    o = Oracle(**cfg)
    s = make_solver(o)
    _, F = o.level_dims(0)
    p_ref, p = np.zeros(F), np.zeros(F)
    ref = o.mg_solve([p_ref], [np.ones(F)])
    got = s.mg_solve(p, np.ones(F))
    assert got["cycles"] == ref["cycles"] == 25
    assert np.array_equal(got["coarse_iters"], ref["coarse_iters"])
    np.testing.assert_allclose(got["residuals"], ref["residuals"], rtol=1e-10, atol=0)     # L2 norms: summation order differs
    assert np.array_equal(p, p_ref)
    dev = s.analytic_deviation()
    assert dev == ref["max_dev"] and dev < 1e-5          # the quadratic is an exact solution of the discrete operator
    s.close()


def test_c2_ldc_257_2d_against_oracle():
    """tests/ldcTest shape at BASELINE's 257^2: PLANAR, depth 4, 7 cycles, Gauss-Seidel, resType 1."""
    cfg = dict(dim=2, xInd=8, zInd=8, vcDepth=4, vcCount=7, gsSmooth=1, resType=1, mgTolerance=1e-6)
    o = Oracle(**cfg)
    s = make_solver(o)
    _, F = o.level_dims(0)
    rhs = smooth_rhs(o, 0, F) + 0.1
    p0 = seeded_field(F, 3, 0.01)
    p_ref = p0.copy()
    ref = o.mg_solve([p_ref], [rhs.copy()])
    s.debug_set_rhs_mean(True, o.last_global_avg())
    p = p0.copy()
    got = s.mg_solve(p, rhs.copy())
    assert got["cycles"] == ref["cycles"]
    assert np.array_equal(got["coarse_iters"], ref["coarse_iters"])
    np.testing.assert_allclose(got["residuals"], ref["residuals"], rtol=1e-10, atol=0)
    assert np.array_equal(p, p_ref)
    s.close()


def test_c3_tgv_257_ppp_against_oracle():
    """BASELINE config 2 shape: 257^3, fully periodic, input/parameters.yaml defaults (depth 3, V(4,4), Gauss-Seidel,
    resType 1); 2 cycles keep the CPU oracle under half a minute."""
    cfg = dict(dim=3, xInd=8, yInd=8, zInd=8, vcDepth=3, vcCount=2, gsSmooth=1, resType=1, preSmooth=4, postSmooth=4,
               xPer=1, yPer=1, zPer=1, mgTolerance=1e-9)
    o = Oracle(**cfg)
    s = make_solver(o)
    _, F = o.level_dims(0)
    rhs = smooth_rhs(o, 0, F)
    p_ref, p = np.zeros(F), np.zeros(F)
    ref = o.mg_solve([p_ref], [rhs.copy()])
    s.debug_set_rhs_mean(True, o.last_global_avg())
    got = s.mg_solve(p, rhs.copy())
    assert got["cycles"] == ref["cycles"]
    assert np.array_equal(got["coarse_iters"], ref["coarse_iters"])
    np.testing.assert_allclose(got["residuals"], ref["residuals"], rtol=1e-10, atol=0)
    assert np.array_equal(p, p_ref)
    s.close()


@pytest.mark.parametrize("gs,mean", [(1, "reference_mean_injected"), (0, "device_mean_vs_exact_mean_oracle")], ids=["gs", "jacobi"])
def test_c4_rbc_513_tanh_against_oracle(gs, mean):
    """BASELINE config 3 shape compared with the oracle DIRECTLY: 513^3, tanh-stretched in all three directions (closed
    box, NNN), depth 6, V(2,2), production mode, 2 cycles (the -O3 IEEE-strict oracle needs under a minute per solve).
    gs: the reference's sequential-sum mean is injected, everything bit-exact.  jacobi: the device forms the mean itself
    (compensated sum over 1.4e8 values) and must match the oracle run with a correctly rounded mean bit for bit."""
    if _host_mem_gb() < 24:
        pytest.skip("needs ~14 GB of host memory")
    cfg = dict(dim=3, xInd=9, yInd=9, zInd=9, vcDepth=6, vcCount=2, gsSmooth=gs, resType=2, mgTolerance=1e-6,
               xGrid=2, yGrid=2, zGrid=2, betaX=1.2, betaY=1.2, betaZ=1.3)
    o = Oracle(fast=True, **cfg)        # -O3, still -ffp-contract=off and no fast-math: same bits as the strict build
    _, F = o.level_dims(0)
    rhs = smooth_rhs(o, 0, F) + 0.1 + 0.01 * seeded_field(F, 11)
    p0 = seeded_field(F, 12, 0.01)
    if mean != "reference_mean_injected":
        o.set_exact_mean(True)
    p_ref = p0.copy()
    ref = o.mg_solve([p_ref], [rhs.copy()])
    avg = o.last_global_avg()
    s = make_solver(o)                  # geometry (metrics, coefficients) is read from the oracle's setup
    o.close()
    if mean == "reference_mean_injected":
        s.debug_set_rhs_mean(True, avg)
    p = p0.copy()
    got = s.mg_solve(p, rhs)
    if mean != "reference_mean_injected":
        assert s.last_rhs_mean() == avg
    assert got["cycles"] == ref["cycles"]
    assert np.array_equal(got["coarse_iters"], ref["coarse_iters"])
    np.testing.assert_allclose(got["residuals"], ref["residuals"], rtol=1e-10, atol=0)
    assert np.array_equal(p, p_ref)
    s.close()


@pytest.mark.parametrize("gs", [0, 1], ids=["jacobi", "gs"])
def test_c4_rbc_513_tanh_scaling_and_repeatability(gs):
    """BASELINE config 3 shape: 513^3, tanh-stretched in all directions (closed box), through the C++ drop-in.
    mgTolerance = 0 switches the two data-dependent tests off (the coarsest solve always runs maxCount iterations, the
    stagnation break never fires), which makes mgSolve exactly homogeneous of degree 1 under scaling by 2."""
    if _host_mem_gb() < 12:
        pytest.skip("needs ~6 GB of host memory")
    cfg = dict(dim=3, xInd=9, yInd=9, zInd=9, vcDepth=6, vcCount=2, gsSmooth=gs, resType=2, mgTolerance=0.0,
               xGrid=2, yGrid=2, zGrid=2, betaX=1.2, betaY=1.2, betaZ=1.3)
    h = host_api.HostSolver(**cfg)
    F = h.F
    rng = np.random.default_rng(4)
    rhs = rng.uniform(-1.0, 1.0, F)
    # the mean of the padded RHS scales exactly too, but it is formed by a parallel sum whose partials could round
    # differently for 2x only through overflow/underflow, which cannot happen here: no injection needed
    p1 = np.zeros(F)
    out1 = h.mg_solve(p1, rhs)
    p1b = np.zeros(F)
    h.mg_solve(p1b, rhs)
    assert np.array_equal(p1, p1b)                      # repeatability
    p2 = np.zeros(F)
    out2 = h.mg_solve(p2, 2.0 * rhs)
    assert out1["cycles"] == out2["cycles"] == 2
    assert np.array_equal(out1["coarse_iters"], out2["coarse_iters"])
    assert np.array_equal(p2, 2.0 * p1)                 # exact homogeneity
    assert np.array_equal(out1["residuals"], out2["residuals"])     # relative norms: the factor cancels exactly
    assert np.isfinite(p1).all() and np.abs(p1).max() > 0
    h.close()


def test_c5_1025_two_kernel_families_and_scaling(monkeypatch):
    """BASELINE config 4 at full size: 1025^3, uniform Neumann box, depth 7, V(2,2) Jacobi (the bench workload).
    (1) The TMA-staged kernels (production path at this size) and the register-marching kernels (verified against the
    oracle at every size the oracle can run) are independent implementations of the same arithmetic: their pressure
    fields must agree bit for bit.  (2) Exact homogeneity under scaling by 2, as in C4."""
    if _host_mem_gb() < 60:
        pytest.skip("needs ~45 GB of host memory")
    cfg = dict(dim=3, xInd=10, yInd=10, zInd=10, vcDepth=7, vcCount=2, gsSmooth=0, resType=2, mgTolerance=0.0)
    rng = np.random.default_rng(5)
    h = host_api.HostSolver(**cfg)
    F = h.F
    Pp, rhs = h.field(0), h.field(1)                    # numpy views of the plainsf storage
    rng.random(out=rhs.reshape(-1))
    rhs -= 0.5
    Pp.fill(0.0)
    assert h.mg_solve_inplace() == 2
    p_tma = Pp.copy()
    assert np.isfinite(p_tma).all() and np.abs(p_tma).max() > 0
    rhs *= 2.0
    Pp.fill(0.0)
    assert h.mg_solve_inplace() == 2
    assert np.array_equal(Pp, 2.0 * p_tma)              # exact homogeneity
    rhs *= 0.5
    rhs_keep = rhs.copy()
    h.close()
    monkeypatch.setenv("SARASMG_NO_TMA", "1")
    h = host_api.HostSolver(**cfg)
    Pp, rhs = h.field(0), h.field(1)
    rhs[...] = rhs_keep
    del rhs_keep
    Pp.fill(0.0)
    assert h.mg_solve_inplace() == 2
    assert np.array_equal(Pp, p_tma)
    h.close()
