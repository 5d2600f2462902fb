"""Times the level-0 Gauss-Seidel sweep alone (CUDA events around smooth()):  python scripts/gs_time.py INDEX [SWEEPS]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SARASMG_PROFILE"] = "1"
os.environ.setdefault("SARASMG_GS_STRIP", "1")
from saras_b200 import capi, host_api  # noqa: E402

index = int(sys.argv[1]) if len(sys.argv) > 1 else 9
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = dict(dim=3, xInd=index, yInd=index, zInd=index, vcDepth=2, vcCount=1, gsSmooth=1, resType=2, mgTolerance=1e-6)
h = host_api.HostSolver(**cfg)
s = h.native
rhs = h.field(1)
np.random.default_rng(1).random(out=rhs.reshape(-1))
s.level_set(0, capi.RESIDUAL, rhs)
s.level_set(0, capi.PRESSURE, rhs)
s.set_level(0)
s.smooth(1)
s.reset_counters()
s.smooth(sweeps)
ms, n = s.profile_query(capi.K_SMOOTH, 0)
n0 = s.level_dims(0)[0]
pts = n0[0] * n0[1] * n0[2]
print("GS_TIME index %d lib %s: %.3f ms per sweep, %.1f GB/s, %.1f cycles-equivalent" % (
    index, os.path.basename(os.environ.get("SARASMG_LIB", "default")), ms / n, 24.0 * pts / (ms / n * 1e-3) / 1e9, 0.0), flush=True)
h.close()
