"""Multi-rank worker, launched by tests/test_multirank.py as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mp_worker.py MODE

MODE = host : CPU, gloo.  One process per rank builds parser -> parallel -> grid -> multigrid_dN (the C++ drop-in's host
              side) in process-rank mode with its MPI collectives riding on torch.distributed, and compares the rank's
              topology, level geometry, pow() coefficients and strided metric copies with the oracle's rank `r`.
MODE = gpu  : one process per GPU.  Every operator of the CUDA path that exchanges halos (updatePads, imposeBC, smooth,
              computeResidual + coarsen, prolong, solve, computeError, vCycle, mgSolve) runs on N GPUs with the NCCL
              face exchange and is compared, rank by rank and bit for bit, with the oracle's emulation of the same
              npX x npY reference run.  The oracle is the checker; the product path never touches it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import mgoracle as mo  # noqa: E402
from oracle.mgoracle import Oracle  # noqa: E402

LAYOUTS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}


def host_cases(world):
    npx, npy = LAYOUTS[world]
    big = 5 if world <= 4 else 6        # local intervals per direction must stay >= 2^(vcDepth+1) (parser.cc:307-337)
    return [
        dict(dim=3, xInd=big, yInd=5, zInd=4, vcDepth=2, npX=npx, npY=npy),
        dict(dim=3, xInd=big, yInd=5, zInd=5, vcDepth=2, npX=npx, npY=npy, xPer=1, yPer=1, zPer=1, xGrid=2, zGrid=2, betaX=1.2, betaZ=1.4, wall_null=0),
        dict(dim=3, xInd=big, yInd=5, zInd=4, vcDepth=2, npX=npx, npY=npy, test_poisson=1),
        dict(dim=2, xInd=big + 1, zInd=5, vcDepth=3, npX=world),
    ]


def run_host(rank, world):
    from saras_b200 import host_api
    for cfg in host_cases(world):
        host_api.install_torch_distributed_hooks(cfg["dim"], cfg.get("test_poisson", 0), rank, world)
        o = Oracle(**cfg)
        h = host_api.HostSolver(**cfg)
        assert h.F == o.level_dims(0)[1]
        assert h.rank_info() == o.rank_info(rank), (h.rank_info(), o.rank_info(rank))
        planar = cfg["dim"] == 2
        used = ["hx", "hz", "hx2", "hz2", "hzhx"] if planar else ["hx", "hy", "hz", "hzhx", "hxhy", "hyhz", "hxhyhz"]
        for q, name in enumerate(mo.COEF_NAMES):
            if name in used:
                assert np.array_equal(h.setup_array(q), np.array([o.coef(l, q) for l in range(o.depth + 1)])), name
        for q, name in enumerate(mo.METRIC_NAMES):
            if planar and q // 2 == 1:
                continue
            ref = np.concatenate([o.metric(rank, l, q) for l in range(o.depth + 1)])
            assert np.array_equal(h.setup_array(10 + q), ref), name
        if cfg.get("test_poisson"):
            for w in range(3):
                ref = o.wall(rank, w).ravel()
                if ref.size:
                    assert np.array_equal(h.setup_array(20 + w), ref), "wall %d" % w
        for d in range(3):
            if not (planar and d == 1):
                assert np.array_equal(h.staggr(d), o.staggr(rank, d))
        h.close()
        dist.barrier()


def gpu_cases(world):
    npx, npy = LAYOUTS[world]
    big = 5 if world <= 4 else 6
    return [
        dict(name="nnn", dim=3, xInd=big, yInd=5, zInd=4, vcDepth=2, npX=npx, npY=npy),
        dict(name="ppp_tanhz", dim=3, xInd=big, yInd=5, zInd=5, vcDepth=2, npX=npx, npY=npy, xPer=1, yPer=1, zPer=1, wall_null=0),
        dict(name="ppn_tanh", dim=3, xInd=big, yInd=5, zInd=4, vcDepth=2, npX=npx, npY=npy, xPer=1, yPer=1, zGrid=2, betaZ=1.5),
        dict(name="testpoisson", dim=3, xInd=big, yInd=5, zInd=4, vcDepth=2, npX=npx, npY=npy, test_poisson=1),
        dict(name="d2_nn", dim=2, xInd=big + 1, zInd=5, vcDepth=3, npX=world),
        dict(name="d2_pp", dim=2, xInd=big + 1, zInd=5, vcDepth=3, npX=world, xPer=1, zPer=1, wall_null=0),
    ]


def run_gpu(rank, world, local):
    from saras_b200 import capi
    from tests.helpers import make_solver, seeded_field, smooth_rhs, rel_linf, check_natural_mean_field

    def bcast_id():
        box = [capi.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def load_both(o, s, level, which, seed, scale=1.0):
        _, F = o.level_dims(level)
        for r in range(world):
            a = seeded_field(F, seed + 101 * r, scale)
            if o.cfg["dim"] == 2:
                a[:, 0, :] = 0.0
                a[:, 2, :] = 0.0
            o.array(r, level, which)[...] = a
            if r == rank:
                s.level_set(level, which, a)

    def same(o, s, level, which, what):
        got = s.level_get(level, which)
        ref = o.array(rank, level, which)
        assert np.array_equal(got, ref), "%s: rank %d level %d array %d differs, max |d| = %g" % (
            what, rank, level, which, float(np.max(np.abs(got - ref))))

    cases = gpu_cases(world)
    ncore = len(cases)
    # the first case once more with the host-driven coarsest solve (one halo exchange + all-gather per iteration)
    cases.append(dict(cases[0], name=cases[0]["name"] + "_coarse_legacy", _env="SARASMG_COARSE_LEGACY=1"))
    # the default halo path is peer memory over CUDA IPC; run two cases with the grouped ncclSend/ncclRecv path as well
    cases.append(dict(cases[0], name=cases[0]["name"] + "_halo_nccl", _env="SARASMG_HALO=nccl"))
    cases.append(dict(cases[1], name=cases[1]["name"] + "_halo_nccl", _env="SARASMG_HALO=nccl"))
    # small levels use the single-CTA peer-memory exchange; force the three-kernel version (large levels) onto these grids too
    cases.append(dict(cases[0], name=cases[0]["name"] + "_ipc_3kernel", _env="SARASMG_IPC_CTA_MAX=0"))
    cases.append(dict(cases[1], name=cases[1]["name"] + "_ipc_3kernel", _env="SARASMG_IPC_CTA_MAX=0"))
    cases.append(dict(cases[3], name=cases[3]["name"] + "_ipc_3kernel", _env="SARASMG_IPC_CTA_MAX=0"))
    # the V-cycle is replayed as a CUDA graph on every rank by default (peer-memory exchanges and the coarsest-level gathers
    # inside); one case with eager launches as well
    cases.append(dict(cases[0], name=cases[0]["name"] + "_nograph", _env="SARASMG_GRAPH_MULTI=0"))
    # the Jacobi sweeps of mgSolve / vCycle push their boundary planes into the neighbours' pads (peer memory); the small
    # grids above run the register-marching sweep: force the TMA-staged one too, and run one case with the push switched off
    cases.append(dict(cases[0], name=cases[0]["name"] + "_push_tma", _env="SARASMG_TMA_MIN=3"))
    cases.append(dict(cases[1], name=cases[1]["name"] + "_push_tma", _env="SARASMG_TMA_MIN=3"))
    cases.append(dict(cases[2], name=cases[2]["name"] + "_nopush", _env="SARASMG_NO_PUSH=1"))
    if os.environ.get("MP_CASES"):      # comma-separated case names (debugging on a budget)
        cases = [c for c in cases if c["name"] in os.environ["MP_CASES"].split(",")]
    if world >= 8:      # an 8-GPU box is charged eight-fold: the six core cases plus one of each variant
        cases = cases[:ncore] + [c for c in cases[ncore:] if c["name"].startswith("nnn") or c["name"].endswith("_nopush")]
    for case in cases:
        base = {k: v for k, v in case.items() if k not in ("name", "_env")}
        for key in ("SARASMG_COARSE_LEGACY", "SARASMG_HALO", "SARASMG_IPC_CTA_MAX", "SARASMG_GRAPH_MULTI", "SARASMG_TMA_MIN", "SARASMG_NO_PUSH"):
            if not os.environ.get("MP_KEEP_ENV"):       # debugging: keep what the caller exported
                os.environ.pop(key, None)
        if case.get("_env"):
            key, val = case["_env"].split("=")
            os.environ[key] = val
        for gs in (0, 1):
            cfg = dict(base, gsSmooth=gs, vcCount=6, resType=2, mgTolerance=1e-7)
            o = Oracle(**cfg)
            assert o.nranks == world
            s = make_solver(o, rank=rank, device=local, nccl_id=bcast_id())
            what = "%s/%s" % (case["name"], "gs" if gs else "jacobi")
            D = o.depth
            # updatePads + imposeBC + smooth at every level
            for level in range(D + 1):
                zb = 1 if level > 0 else 0
                load_both(o, s, level, mo.PRESSURE, 10 + level)
                load_both(o, s, level, mo.RESIDUAL, 20 + level, 5.0)
                o.set_level(level); o.set_zero_bc(zb)
                s.set_level(level); s.set_zero_bc(zb)
                o.update_pads(mo.RESIDUAL); s.update_pads(capi.RESIDUAL)
                same(o, s, level, mo.RESIDUAL, what + " updatePads")
                o.impose_bc(); s.impose_bc()
                same(o, s, level, mo.PRESSURE, what + " imposeBC")
                for count in (1, 2):
                    o.smooth(count); s.smooth(count)
                    same(o, s, level, mo.PRESSURE, what + " smooth(%d)" % count)
            # residual -> restriction (reads edge pads) -> prolongation
            for level in range(D):
                # fresh arrays per level: the oracle's Jacobi swap exchanges P and TMP of ALL levels at once (d3:168), so
                # after the smoothing block above its other levels' names point at stale buffers (dead data in a real cycle)
                load_both(o, s, level, mo.PRESSURE, 30 + level)
                load_both(o, s, level, mo.RESIDUAL, 50 + level, 3.0)
                o.set_level(level); s.set_level(level)
                o.set_zero_bc(1); s.set_zero_bc(1)
                o.compute_residual(); s.compute_residual()
                same(o, s, level, mo.TMP, what + " computeResidual")
                o.coarsen(); s.coarsen()
                same(o, s, level + 1, mo.RESIDUAL, what + " coarsen")
            # coarsest solve with its global MAX test
            load_both(o, s, D, mo.PRESSURE, 77, 0.0)
            load_both(o, s, D, mo.RESIDUAL, 78, 0.01)
            o.set_level(D); s.set_level(D)
            it_ref = o.solve_coarsest(); it_got = s.solve_coarsest()
            assert it_ref == it_got, (what, it_ref, it_got)
            same(o, s, D, mo.PRESSURE, what + " solve")
            # whole V-cycle
            load_both(o, s, 0, mo.PRESSURE, 31, 0.1)
            load_both(o, s, 0, mo.RESIDUAL, 32, 1.0)
            for level in range(1, D + 1):
                load_both(o, s, level, mo.PRESSURE, 40 + level, 0.0)
            o.vcycle(); s.vcycle()
            same(o, s, 0, mo.PRESSURE, what + " vCycle")
            for norm in (1, 2):
                o.set_level(0); s.set_level(0)
                a, b = o.compute_error(norm), s.compute_error(norm)
                assert abs(a - b) <= 1e-13 * abs(a), (what, norm, a, b)
            # mgSolve
            _, F = o.level_dims(0)
            if cfg.get("test_poisson"):
                rhs = [np.ones(F) for _ in range(world)]
                p0 = [np.zeros(F) for _ in range(world)]
            else:
                rhs = [smooth_rhs(o, r, F) + 0.3 for r in range(world)]
                p0 = [seeded_field(F, 7 + r, 0.01) for r in range(world)]
            p_ref = [a.copy() for a in p0]
            ref = o.mg_solve(p_ref, [a.copy() for a in rhs])
            modes = ["natural"] if cfg.get("test_poisson") else ["injected_mean", "natural"]
            if os.environ.get("MP_NATURAL_FIRST"):
                modes = modes[::-1]
            for mode in modes:
                s.debug_set_rhs_mean(mode == "injected_mean", o.last_global_avg())
                p = p0[rank].copy()
                got = s.mg_solve(p, rhs[rank].copy())
                assert got["cycles"] == ref["cycles"], (what, got["cycles"], ref["cycles"])
                assert np.array_equal(got["coarse_iters"], ref["coarse_iters"]), (what, got["coarse_iters"], ref["coarse_iters"])
                np.testing.assert_allclose(got["residuals"], ref["residuals"], rtol=1e-10, atol=0)
                if mode == "natural" and not cfg.get("test_poisson"):
                    check_natural_mean_field(o, p0, rhs, rank, p, p_ref, what, got_mean=s.last_rhs_mean())
                else:
                    assert np.array_equal(p, p_ref[rank]), (what, mode, float(np.max(np.abs(p - p_ref[rank]))))
            # device-resident projection step across ranks: divergence -> mgSolve -> Pp.syncData (faces with
            # faceRanks, edges with edgeRanks over grouped ncclSend/ncclRecv) -> P += Pp -> V -= dt grad(Pp), two time steps
            if cfg["dim"] == 3 and not cfg.get("test_poisson"):
                n, _ = o.level_dims(0)
                sh = mo.proj_shapes(n)
                met = [mo.proj_metrics(o, r) for r in range(world)]
                rng = np.random.default_rng(4242)
                fld = [[np.ascontiguousarray(rng.uniform(-1.0, 1.0, sh[k])) for k in ("vx", "vy", "vz", "sc")] + [np.zeros(sh["sc"])] for _ in range(world)]
                s.debug_set_rhs_mean(False)
                o.set_exact_mean(True)
                s.proj_setup(*met[rank])
                for which, a in zip((capi.PROJ_VX, capi.PROJ_VY, capi.PROJ_VZ, capi.PROJ_P, capi.PROJ_PP), fld[rank]):
                    s.proj_upload(which, a)
                dt = 0.01
                for step in range(2):
                    got = s.proj_step(dt)
                    rhs_l = [mo.proj_divergence(n, f[0], f[1], f[2], met[r][0], met[r][2], 1.0 / dt) for r, f in enumerate(fld)]
                    pp_l = [f[4] for f in fld]
                    refp = o.mg_solve(pp_l, rhs_l)
                    mo.proj_sync_multi(o, pp_l)
                    for r, f in enumerate(fld):
                        mo.proj_correct_nosync(n, f[4], f[3], f[0], f[1], f[2], met[r][1], met[r][2], dt)
                    assert got["cycles"] == refp["cycles"], (what, step, got["cycles"], refp["cycles"])
                    np.testing.assert_allclose(got["residuals"], refp["residuals"], rtol=1e-10, atol=0)
                    for which, ref_a, nm in zip((capi.PROJ_VX, capi.PROJ_VY, capi.PROJ_VZ, capi.PROJ_P, capi.PROJ_PP), fld[rank], ("Vx", "Vy", "Vz", "P", "Pp")):
                        g = s.proj_download(which)
                        assert np.array_equal(g, ref_a), "%s projection step %d: %s differs on rank %d, max |d| = %g" % (what, step, nm, rank, float(np.max(np.abs(g - ref_a))))
                o.set_exact_mean(False)
            assert s.kernel_launches() > 0
            s.close()
            dist.barrier()
        if rank == 0:
            print("mp_worker: case %s ok on %d ranks" % (case["name"], world), flush=True)


def main():
    mode = sys.argv[1]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")           # control plane only; the halo traffic is the library's own NCCL communicator
    try:
        if mode == "host":
            run_host(rank, world)
        else:
            if not torch.cuda.is_available() or torch.cuda.device_count() < world:
                raise SystemExit("mp_worker gpu: needs %d CUDA devices" % world)
            run_gpu(rank, world, local)
        dist.barrier()
        if rank == 0:
            print("MP_WORKER_OK %s %d" % (mode, world), flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
