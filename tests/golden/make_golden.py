"""Generates tests/golden/*.npz from the REFERENCE ITSELF: oracle/_ref = /root/reference's lib/poisson, lib/grid,
lib/parallel, lib/io/parser sources compiled unmodified against stand-in headers (oracle/refbuild/Makefile).

    python -m tests.golden.make_golden        (only where /root/reference exists; the fixtures are committed)

Each fixture holds, per case: the V-cycle count, the printed residual history (17 significant digits in production mode,
16 under TEST_POISSON as the reference prints it), the final padded pressure field of every rank, and the reference's
level-0 metric arrays and per-level coefficients.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import sarasref  # noqa: E402
from oracle.mgoracle import Oracle  # noqa: E402
from tests.golden.cases import GOLDEN_CASES, inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sarasref.build()
    for name, cfg in GOLDEN_CASES:
        o = Oracle(**cfg)          # used for extents and input generation only
        p, rhs = inputs(o, cfg)
        res = sarasref.mg_solve(p, [r.copy() for r in rhs], **cfg)
        met, coef = sarasref.setup_data(**cfg)
        out = dict(cycles=np.int64(res["cycles"]), residuals=res["residuals"], max_dev=np.float64(res["max_dev"]), coef=coef)
        for r, a in enumerate(p):
            out["p_rank%d" % r] = a
        for r, d in enumerate(met):
            for k, v in d.items():
                out["%s_rank%d" % (k, r)] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-28s cycles=%d last residual=%.6e" % (name, res["cycles"], res["residuals"][-1]))


if __name__ == "__main__":
    main()
