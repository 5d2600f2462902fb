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
