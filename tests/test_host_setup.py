"""CPU: the host-side logic of the drop-in (saras_b200/host/poisson.cc + mirror grid/parallel/parser): level sizes, pow()
coefficients, strided metric copies, Dirichlet planes and rank topology must equal the oracle's restatement of the
reference bit for bit (the oracle itself is pinned to the reference by tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from oracle import mgoracle as mo
from oracle.mgoracle import Oracle
from saras_b200 import host_api

CASES = [
    dict(dim=3, xInd=4, yInd=4, zInd=4, vcDepth=2),
    dict(dim=3, xInd=5, yInd=4, zInd=6, vcDepth=3, xGrid=2, yGrid=2, zGrid=2, betaX=1.1, betaY=1.3, betaZ=1.7, Lx=2.0, Ly=1.0, Lz=1.5),
    dict(dim=3, xInd=7, yInd=7, zInd=7, vcDepth=5, test_poisson=1),
    dict(dim=3, xInd=5, yInd=5, zInd=5, vcDepth=3, xPer=1, yPer=1, zPer=0, zGrid=2, betaZ=1.2),
    dict(dim=2, xInd=7, zInd=7, vcDepth=4),
    dict(dim=2, xInd=6, zInd=5, vcDepth=3, test_poisson=1, Lx=2.0, Lz=0.5),
    dict(dim=2, xInd=5, zInd=6, vcDepth=3, xPer=1, zPer=1, xGrid=2, zGrid=2, betaX=1.4, betaZ=0.9),
]


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: "d%d_%d%d%d%s" % (c["dim"], c["xInd"], c.get("yInd", 0), c["zInd"], "_t" if c.get("test_poisson") else ""))
def test_host_setup_bitexact(cfg):
    o = Oracle(**cfg)
    h = host_api.HostSolver(**cfg)
    _, F = o.level_dims(0)
    assert h.F == F
    assert h.rank_info() == o.rank_info(0)
    planar = cfg["dim"] == 2
    used3 = ["hx", "hy", "hz", "hzhx", "hxhy", "hyhz", "hxhyhz"]
    used2 = ["hx", "hz", "hx2", "hz2", "hzhx"]
    for q, name in enumerate(mo.COEF_NAMES):
        if name in (used2 if planar else used3):
            ref = np.array([o.coef(l, q) for l in range(o.depth + 1)])
            assert np.array_equal(h.setup_array(q), ref), name
    for q, name in enumerate(mo.METRIC_NAMES):
        if planar and q // 2 == 1:
            continue
        ref = np.concatenate([o.metric(0, l, q) for l in range(o.depth + 1)])
        assert np.array_equal(h.setup_array(10 + q), ref), name
    if cfg.get("test_poisson"):
        for w in range(3):
            ref = o.wall(0, w).ravel()
            if ref.size:
                assert np.array_equal(h.setup_array(20 + w), ref), "wall %d" % w
    assert int(h.setup_array(30)[0]) == o.max_count
    for d in range(3):
        if planar and d == 1:
            continue
        assert np.array_equal(h.staggr(d), o.staggr(0, d))
    h.close()


def test_parser_rejects_bad_depth():
    with pytest.raises(RuntimeError) as e:
        host_api.HostSolver(dim=3, xInd=4, yInd=4, zInd=4, vcDepth=4)      # 16 intervals < 2^(4+1)
    assert "too coarse" in str(e.value)


def test_yaml_parser_reads_reference_style_file(tmp_path):
    """the mgtest driver reads the reference's own parameters.yaml format (tests/mgTest/input/parameters.yaml)"""
    import subprocess, os
    exe = os.path.join(os.path.dirname(host_api.__file__), "lib", "mgtest_d3")
    assert os.path.exists(exe)
    y = tmp_path / "parameters.yaml"
    from tests.helpers import parameters_yaml
    y.write_text(parameters_yaml(dim=3, xInd=4, yInd=4, zInd=4, vcDepth=2, vcCount=3))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except ImportError:
        has_gpu = False
    r = subprocess.run([exe, str(y)], capture_output=True, text=True)
    if has_gpu:
        assert "Residual after V Cycle 2 is" in r.stdout
        assert "Maximum absolute deviation from analytic solution is:" in r.stdout
    else:
        # YAML parsed, objects built, then the device layer refuses loudly: no CPU fallback
        assert r.returncode != 0 and "CUDA" in r.stderr
