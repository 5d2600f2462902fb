"""Runs each multigrid operator a few times on level 0 (and the skewed GS sweep) so that ncu can capture them in isolation.
    ncu --set full --clock-control none --import-source on -k regex:'k_stencil|k_error|k_prolong|k_coarsen|k_gs_skew' -o gpurun_out/prof python scripts/prof_ops.py 9
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saras_b200 import capi, host_api  # noqa: E402

index = int(sys.argv[1]) if len(sys.argv) > 1 else 9
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mesh = sys.argv[3] if len(sys.argv) > 3 else "uniform"
for gs in (0, 1):
    cfg = dict(dim=3, xInd=index, yInd=index, zInd=index, vcDepth=3, vcCount=1, gsSmooth=gs, resType=2, mgTolerance=1e-6)
    if mesh == "tanh":
        cfg.update(zGrid=2, betaZ=1.3)
    h = host_api.HostSolver(**cfg)
    s = h.native
    rhs = h.field(1)
    rng = np.random.default_rng(1)
    rng.random(out=rhs.reshape(-1))
    s.level_set(0, capi.RESIDUAL, rhs)
    s.level_set(0, capi.PRESSURE, rhs * 0.5)
    for _ in range(reps):
        s.set_level(0)
        s.smooth(2)
        if gs == 0:
            s.compute_residual()
            s.coarsen()
            s.set_level(1)
            s.prolong()
            s.compute_error(2)
    h.close()
print("done")
