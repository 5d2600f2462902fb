"""Shared test helpers: build the CUDA solver of one rank from the same geometry the oracle uses,
seeded inputs, and comparison utilities.  The oracle is the CHECKER here, never the thing under test."""
import numpy as np

from oracle import mgoracle as mo
from saras_b200 import capi


def geometry_from_oracle(o, rank=0):
    """Everything sarasmg_desc needs, taken from the oracle's restatement of grid.cc / poisson.cc setup
    (inputs only: the arithmetic under test happens on the GPU)."""
    D = o.depth
    n0, _ = o.level_dims(0)
    coefs = {name: np.array([o.coef(l, i) for l in range(D + 1)]) for i, name in enumerate(mo.COEF_NAMES)}
    metrics = {name: np.concatenate([o.metric(rank, l, i) for l in range(D + 1)]) for i, name in enumerate(mo.METRIC_NAMES)}
    walls = dict(x_wall=o.wall(rank, 0), y_wall=o.wall(rank, 1), z_wall=o.wall(rank, 2))
    info = o.rank_info(rank)
    return dict(n0=n0, coefs=coefs, metrics=metrics, walls=walls, info=info)


def make_solver(o, rank=0, device=0, nccl_id=None, flags=0, print_residual=0):
    c = o.cfg
    g = geometry_from_oracle(o, rank)
    return capi.Solver(
        dim=c["dim"], vc_depth=c["vcDepth"], vc_count=c["vcCount"], pre_smooth=c["preSmooth"], post_smooth=c["postSmooth"],
        gs_smooth=c["gsSmooth"], res_type=c["resType"], mg_tolerance=c["mgTolerance"],
        x_per=c["xPer"], y_per=c["yPer"], z_per=c["zPer"], n0=g["n0"], coefs=g["coefs"], metrics=g["metrics"],
        walls=g["walls"] if c["test_poisson"] else None, test_poisson=c["test_poisson"], print_residual=print_residual,
        rank=rank, n_proc=o.nranks, npx=c["npX"], npy=c["npY"], x_rank=g["info"]["xRank"], y_rank=g["info"]["yRank"],
        face_ranks=g["info"]["faceRanks"], device=device, nccl_id=nccl_id, flags=flags)


def seeded_field(F, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.uniform(-1.0, 1.0, size=F) * scale)


def smooth_rhs(o, rank, F):
    """zero-mean-ish analytic RHS sampled on the rank's staggered coordinates (production-mode cases)."""
    x = o.staggr(rank, 0)[:, None, None]
    y = o.staggr(rank, 1)[None, :, None] if o.cfg["dim"] == 3 else np.zeros((1, F[1], 1))
    z = o.staggr(rank, 2)[None, None, :]
    c = o.cfg
    r = np.cos(2 * np.pi * x / c["Lx"]) * np.cos(2 * np.pi * y / (c["Ly"] if c["dim"] == 3 else 1.0)) * np.cos(4 * np.pi * z / c["Lz"])
    return np.ascontiguousarray(np.broadcast_to(r, F).copy())


def max_abs(a):
    return float(np.max(np.abs(a))) if a.size else 0.0


def rel_linf(a, b):
    den = max_abs(b)
    return max_abs(a - b) / den if den else max_abs(a - b)


def check_natural_mean_field(o, p0_list, rhs_list, rank, p_got, p_ref_seq, what="", got_mean=None):
    """Production mode with the DEVICE-computed RHS mean (poisson.cc:122-128).

    The reference forms the mean with one sequential fp64 sum; its rounding error (~sqrt(n) ulp, traversal-order
    dependent) cannot be reproduced by a parallel reduction.  The device forms the correctly rounded sum instead, so
      (1) against the oracle run with a correctly rounded mean (Oracle.set_exact_mean: everything else is the reference's
          arithmetic) the field must be BIT-EXACT -- the whole mgSolve path, mean included, is pinned;
      (2) against the reference's sequential-sum run the field must be within north_star's 1e-12 relative L-inf, unless the
          reference's own sum is visibly off the exact one, in which case the deviation is the reference's rounding error
          amplified by the singular Neumann / periodic solve (measured: <= 4e-12 on every case of this suite) and is
          bounded at 1e-11."""
    seq_avg = o.last_global_avg()
    o.set_exact_mean(True)
    pe = [a.copy() for a in p0_list]
    o.mg_solve(pe, [a.copy() for a in rhs_list])
    exact_avg = o.last_global_avg()
    o.set_exact_mean(False)
    assert np.array_equal(p_got, pe[rank]), "%s: natural mode differs from the exact-mean oracle by %g (device mean %r, exact-mean oracle %r, sequential-sum oracle %r)" % (
        what, max_abs(p_got - pe[rank]), got_mean, exact_avg, seq_avg)
    dev = rel_linf(p_got, p_ref_seq[rank])
    assert dev <= 1e-12 or (seq_avg != exact_avg and dev <= 1e-11), (what, dev, seq_avg, exact_avg)
    return dev


def parameters_yaml(dim=3, xInd=4, yInd=4, zInd=4, npX=1, npY=1, domain="NNN", mesh="UUU", beta=(1.0, 1.0, 1.0), length=(1.0, 1.0, 1.0),
                    vcDepth=2, vcCount=3, tol=1.0e-6, gs=1, pre=2, post=2, resType=2, print_residual=True):
    """input/parameters.yaml text with the key set lib/io/parser.cc:69-239 reads (sections Program / Mesh / Parallel /
    Solver / Multigrid), including a comment, a trailing comment and the folded '>' Probes scalar the parser must cope with."""
    sections = [
        ("Program", [("Problem Type", 2), ("Initial Condition", 0), ("Domain Type", '"%s"      # periodicity per axis' % domain), ("RBC Type", 1),
                     ("Reynolds Number", 1000), ("Rossby Number", 12), ("Rayleigh Number", 10000), ("Prandtl Number", 0.71),
                     ("Taylor Number", 0.0), ("X Length", length[0]), ("Y Length", length[1]), ("Z Length", length[2]),
                     ("Heating Plate", "false"), ("Plate Radius", 0.25), ("Force", 0)]),
        ("Mesh", [("Mesh Type", '"%s"' % mesh), ("X Beta", beta[0]), ("Y Beta", beta[1]), ("Z Beta", beta[2]),
                  ("X Index", xInd), ("Y Index", 0 if dim == 2 else yInd), ("Z Index", zInd)]),
        ("Parallel", [("Number of OMP threads", 1), ("X Number of Procs", npX), ("Y Number of Procs", npY)]),
        ("Solver", [("Differentiation Scheme", 1), ("Integration Scheme", 1), ("Solve Tolerance", 1.0e-6), ("Restart Run", "false"),
                    ("Use CFL Condition", "false"), ("Courant Number", 0.1), ("Time-Step", 0.001), ("Final Time", 0.001), ("I/O Count", 1),
                    ("Solution Format", 1), ("Solution Write Interval", 0.001), ("Restart Write Interval", 0.1), ("Record Probes", "false"),
                    ("Probe Time Interval", 0.1), ("Probes", ">\n        [5, 5, 5]\n        [1:62:3, 7, 1:62:3]")]),
        ("Multigrid", [("V-Cycle Depth", vcDepth), ("V-Cycle Count", vcCount), ("Solve Tolerance", tol), ("Smoothing Method", gs),
                       ("Pre-Smoothing Count", pre), ("Post-Smoothing Count", post), ("Residual Type", resType),
                       ("Print Residual", "true" if print_residual else "false")]),
    ]
    out = ["# generated by tests/helpers.py"]
    for name, items in sections:
        out.append('"%s":' % name)
        for k, v in items:
            out.append('    "%s": %s' % (k, v))
    return "\n".join(out) + "\n"
