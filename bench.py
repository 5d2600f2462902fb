#!/usr/bin/env python
"""bench.py -- multigrid V-cycle throughput of the pressure-Poisson hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3                    # this repo's CUDA path (default)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the reference's own CPU code on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                       # N = 2, 4, 8: x/y pencils, NCCL halos

Workload: 3-D Poisson problem on the 1025^3 grid (Index 10), Neumann walls (NNN), uniform
mesh, V(2,2) cycles of depth 7, at most `--cycles` (10) V-cycles per mgSolve at the reference's default
mgTolerance 1e-6 (the random RHS never stagnates within 10 cycles; the executed count is reported), Jacobi smoother (`--smoother gs` selects the reference YAMLs' lexicographic Gauss-Seidel), RHS = seeded
uniform(-1,1) noise.  The SAME global problem is split npX x npY = 1x1, 2x1, 2x2, 4x2 over 1, 2, 4, 8 GPUs (strong scaling).

A "step" is one poisson::mgSolve.  `value` times `--steps` solves with Pp / mgRHS already resident in HBM
(sarasmg_solve_resident); `e2e` times the drop-in C++ call mgSolver.mgSolve(Pp, mgRHS) on HOST plainsf arrays, the
host->device and device->host copies inside the timed region.  LUP = one stencil evaluation at one core grid point
(smoother update, residual, error norm, coarsest-level GS iteration): sum_l N_l (pre+post+1) + N_0 per V-cycle with the
coarsest level counted by its actual iteration count.

The oracle (oracle/) is used here only for the cpu_baseline / --impl reference legs: oracle/_ref is the reference's own
lib/poisson compiled unmodified with its -Ofast flags, run with one thread-rank per core.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries the one JSON line; if the environment asks NCCL for debug output (NCCL_DEBUG=VERSION prints a banner),
# send it to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "multigrid_vcycle_glups"
UNIT = "GLUP/s"
LAYOUTS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}
BYTES_PER_LUP = {"smooth": 24.0, "residual": 24.0, "coarsen": 9.0, "prolong": 17.0, "error": 16.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--index", type=int, default=10, help="grid index: 2^index + 1 points per direction (10 -> 1025^3)")
    ap.add_argument("--depth", type=int, default=-1, help="V-cycle depth (default: index - 3, the C5 setting 7 at index 10)")
    ap.add_argument("--cycles", type=int, default=10, help="V-cycles per mgSolve")
    ap.add_argument("--smoother", default="jacobi", choices=["jacobi", "gs"])
    ap.add_argument("--mesh", default="uniform", choices=["uniform", "tanh"])
    ap.add_argument("--tolerance", type=float, default=1.0e-6, help="parser::mgTolerance (reference YAML default 1e-6): coarsest-level solve and stagnation test")
    ap.add_argument("--breakdown", action="store_true", help="one extra solve with every level profiled; per-(operator, level) table on stderr")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ttt", action="store_true", help="skip the time-to-tolerance solve")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-index", type=int, default=8, help="grid index of the bounded CPU sample (8 -> 257^3)")
    return ap.parse_args()


def solver_cfg(a, index, npx, npy, depth):
    cfg = dict(dim=3, test_poisson=0, xInd=index, yInd=index, zInd=index, npX=npx, npY=npy, xPer=0, yPer=0, zPer=0,
               vcDepth=depth, vcCount=a.cycles, preSmooth=2, postSmooth=2, gsSmooth=1 if a.smoother == "gs" else 0,
               resType=2, mgTolerance=a.tolerance, wall_null=1)
    if a.mesh == "tanh":
        cfg.update(zGrid=2, betaZ=1.3)
    return cfg


def lups_per_solve(level_points, pre, post, coarse_iters):
    """level_points[l] = core points processed on level l summed over ranks; coarse_iters = per-cycle GS iterations."""
    D = len(level_points) - 1
    per_cycle_fixed = sum(level_points[l] * (pre + post + 1) for l in range(D)) + level_points[0]
    return sum(per_cycle_fixed + level_points[D] * int(it) for it in coarse_iters)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200", "-i", str(device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), power_w_max=float(max(power)),
                    samples=len(sm), reasons=sorted(reasons))


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference leg: oracle/_ref = the reference's own lib/poisson sources, -Ofast, thread-ranks
# ---------------------------------------------------------------------------------------------------------------------
def cpu_layout(cores):
    ranks = 1
    while ranks * 2 <= min(cores, 64):
        ranks *= 2
    npx = 1
    while npx * npx < ranks:
        npx *= 2
    return npx, ranks // npx


def reference_sample(a, cycles, repeats=1, warm=0):
    """Times poisson::mgSolve of the reference on a bounded sample of the bench workload.  Returns (glups, info)."""
    from oracle import sarasref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    kind = "reference" if sarasref.available(3, 0, fast=True) else "port"
    index = a.cpu_index
    depth = index - 3
    npx, npy = cpu_layout(cores) if kind == "reference" else (1, 1)
    cfg = solver_cfg(a, index, npx, npy, depth)
    cfg["vcCount"] = cycles
    nr = npx * npy
    rng = np.random.default_rng(2024)
    if kind == "reference":
        F = sarasref.level0_extents(sarasref._cfg(cfg))
    else:
        from oracle.mgoracle import Oracle
        orc = Oracle(fast=True, **cfg)
        _, F = orc.level_dims(0)
    R = [np.ascontiguousarray(rng.uniform(-1.0, 1.0, F)) for _ in range(nr)]
    n = [F[0] - 2, F[1] - 2, F[2] - 2]
    level_points = [nr * (((n[0] - 1) >> l) + 1) * (((n[1] - 1) >> l) + 1) * (((n[2] - 1) >> l) + 1) for l in range(depth + 1)]
    times = []
    for it in range(warm + repeats):
        P = [np.zeros(F) for _ in range(nr)]
        t0 = time.perf_counter()
        if kind == "reference":
            out = sarasref.mg_solve(P, R, fast=True, **cfg)
        else:
            out = orc.mg_solve(P, R)
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    # the reference does not report its coarsest-level iteration counts; count the coarsest level with `pre+post` sweeps
    coarse = out.get("coarse_iters", [cfg["preSmooth"] + cfg["postSmooth"]] * out["cycles"])
    lups = lups_per_solve(level_points, cfg["preSmooth"], cfg["postSmooth"], coarse)
    sec = float(np.mean(times))
    info = dict(kind=kind, cores=nr, host_cores=cores,
                sample="%d^3 grid (Index %d), %d V(2,2) cycles, depth %d, %s smoother, %dx%d thread-ranks, %s" % (
                    n[2], index, out["cycles"], depth, a.smoother, npx, npy,
                    "reference lib/poisson compiled -Ofast (oracle/_ref)" if kind == "reference" else "oracle/mg_oracle.c -O3"),
                seconds_per_solve=sec, lups_per_solve=lups)
    return lups / sec / 1e9, info, times


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    glups, info, times = reference_sample(a, a.cycles, repeats=a.steps, warm=a.warmup)
    ms = 1e3 * float(np.mean(times))
    npx, npy = LAYOUTS.get(a.gpus, (1, 1))
    line = dict(metric=METRIC, value=glups, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=workload_name(a, npx, npy), sample=info["sample"]),
                cpu_baseline=dict(value=glups, unit=UNIT, cores=info["cores"], kind=info["kind"], sample=info["sample"]),
                e2e=dict(value=glups, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def workload_name(a, npx, npy):
    n = (1 << a.index) + 1
    depth = a.depth if a.depth >= 0 else a.index - 3
    return "C5 Poisson %d^3 (Index %d) NNN %s mesh, %d x V(2,2) depth %d, %s smoother, resType 2, np %dx%d" % (
        n, a.index, a.mesh, a.cycles, depth, "lexicographic Gauss-Seidel" if a.smoother == "gs" else "Jacobi", npx, npy)


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    from saras_b200 import capi, host_api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (a.gpus, world))
    if a.gpus not in LAYOUTS:
        raise SystemExit("bench.py: --gpus must be 1, 2, 4 or 8")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    os.environ["SARASMG_DEVICE"] = str(local)
    os.environ["SARASMG_PROFILE"] = "1"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_api.install_torch_distributed_hooks(3, 0, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    npx, npy = LAYOUTS[a.gpus]
    depth = a.depth if a.depth >= 0 else a.index - 3
    cfg = solver_cfg(a, a.index, npx, npy, depth)
    h = host_api.HostSolver(**cfg)                 # parser -> parallel -> grid -> multigrid_d3 -> plainsf Pp, mgRHS
    Pp, mgRHS = h.field(0), h.field(1)             # numpy views of the plainsf (Blitz) storage
    rng = np.random.default_rng(2024 + rank)
    flat = mgRHS.reshape(-1)
    rng.random(out=flat)
    flat *= 2.0
    flat -= 1.0
    Pp.fill(0.0)
    s = h.native                                    # creates the device hierarchy (collective)
    s.set_profile_max_level(1)
    pre, post = cfg["preSmooth"], cfg["postSmooth"]
    dims = [s.level_dims(l)[0] for l in range(depth + 1)]
    level_points_local = [d[0] * d[1] * d[2] for d in dims]
    level_points = [allsum(v) for v in level_points_local]
    nbytes = Pp.nbytes

    # ---- device-resident arm: inputs already in HBM --------------------------------------------------------------
    s.host_register(Pp)
    s.host_register(mgRHS)
    s.upload(Pp, mgRHS)
    s.synchronize()
    for _ in range(a.warmup):
        out = s.solve_resident()
    s.reset_counters()
    barrier(); s.synchronize(); torch.cuda.synchronize()
    clocks = ClockSampler(local) if rank == 0 else None
    s.event_record(0)
    outs = []
    for _ in range(a.steps):
        outs.append(s.solve_resident())
    s.event_record(1)
    ms_total = s.event_elapsed_ms(0, 1)
    s.synchronize(); torch.cuda.synchronize(); barrier()
    clk = clocks.stop() if clocks else None
    ms_total = allmax(ms_total)
    launches = allsum(s.kernel_launches())
    lups_step = lups_per_solve(level_points, pre, post, outs[-1]["coarse_iters"])
    value = lups_step * a.steps / (ms_total * 1e-3) / 1e9

    # per-operator device times of the fine levels, measured live with CUDA events over the timed region
    peak, peak_src = measured_peak_hbm()
    per_level = []
    for lvl in range(min(2, depth + 1)):
        for kind, name in ((capi.K_SMOOTH, "smooth"), (capi.K_RESIDUAL, "residual"), (capi.K_COARSEN, "coarsen"),
                           (capi.K_PROLONG, "prolong"), (capi.K_ERROR, "error")):
            ms, n = s.profile_query(kind, lvl)
            if n == 0:
                continue
            pts = level_points_local[lvl]
            gbs = BYTES_PER_LUP[name] * pts / (ms / n * 1e-3) / 1e9
            per_level.append(dict(level=lvl, op=name, launches=int(n), avg_ms=ms / n, gbs=gbs, frac=gbs / peak))
    dom = next((e for e in per_level if e["level"] == 0 and e["op"] == "smooth"), None)
    roofline = None
    if dom:
        traffic = None      # DRAM bytes per launch of this kernel from the committed ncu capture (same grid, same layout only)
        if a.gpus == 1 and a.index == 10 and a.mesh == "uniform":
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_1025_n1.json")))[a.smoother]["bytes"]
            except Exception:
                traffic = None
        roofline = dict(bound="hbm", achieved=dom["gbs"], peak=peak, unit="GB/s", frac=dom["gbs"] / peak, traffic=traffic,
                        kernel="level-0 smoother sweep (%s)" % ("k_gs_skew<3> lexicographic Gauss-Seidel" if a.smoother == "gs" else "k_stencil_tma<0> Jacobi, TMA-staged x-march"),
                        algorithmic_bytes_per_launch=BYTES_PER_LUP["smooth"] * level_points_local[0], avg_launch_ms=dom["avg_ms"],
                        peak_source=peak_src)

    breakdown = None
    if a.breakdown:
        s.set_profile_max_level(1 << 20)
        s.reset_counters()
        s.solve_resident()
        breakdown = []
        for kind in range(len(capi.KIND_NAMES)):
            for lvl in range(depth + 1):
                ms, n = s.profile_query(kind, lvl)
                if n:
                    breakdown.append(dict(op=capi.KIND_NAMES[kind], level=lvl, launches=int(n), ms=ms))
        if rank == 0:
            tot = sum(e["ms"] for e in breakdown)
            for e in sorted(breakdown, key=lambda e: -e["ms"])[:40]:
                print("breakdown %-13s level %d  n=%6d  %10.3f ms  %5.1f%%" % (e["op"], e["level"], e["launches"], e["ms"], 100 * e["ms"] / tot), file=sys.stderr)
            print("breakdown total %.3f ms in events, last_solve_ms %.3f" % (tot, s.last_solve_ms()), file=sys.stderr)
        s.set_profile_max_level(1)

    # per-(operator, level) achieved bandwidth of EVERY level: one extra solve outside the timed region with all levels bracketed
    s.set_profile_max_level(1 << 20)
    s.reset_counters()
    s.solve_resident()
    per_level_all = []
    for lvl in range(depth + 1):
        for kind, name in ((capi.K_SMOOTH, "smooth"), (capi.K_RESIDUAL, "residual"), (capi.K_COARSEN, "coarsen"),
                           (capi.K_PROLONG, "prolong"), (capi.K_ERROR, "error"), (capi.K_HALO, "halo"), (capi.K_BC, "bc"),
                           (capi.K_COARSE_SOLVE, "coarse_solve")):
            ms, n = s.profile_query(kind, lvl)
            if n == 0:
                continue
            e = dict(level=lvl, op=name, launches=int(n), avg_us=1e3 * ms / n)
            if name in BYTES_PER_LUP:
                gbs = BYTES_PER_LUP[name] * level_points_local[lvl] / (ms / n * 1e-3) / 1e9
                e.update(gbs=gbs, frac=gbs / peak)
            per_level_all.append(e)
    s.set_profile_max_level(1)

    # time to tolerance: the same hierarchy run until the reference's stagnation test |r_prev - r| < mgTolerance stops it
    ttt = None
    if not a.no_ttt:
        s.set_vc_count(60)
        s.reset_counters()
        barrier(); s.synchronize()
        s.event_record(2)
        out_t = s.solve_resident()
        s.event_record(3)
        ms_t = allmax(s.event_elapsed_ms(2, 3))
        ttt = dict(ms=ms_t, cycles=int(out_t["cycles"]), tolerance=a.tolerance, final_residual=float(out_t["residuals"][-1]),
                   rule="poisson.cc:145 stagnation test, at most 60 cycles")
        s.set_vc_count(a.cycles)

    # ---- end-to-end arm: mgSolver.mgSolve(Pp, mgRHS) on host plainsf arrays ---------------------------------------
    e2e = None
    if not a.no_e2e:
        e2e_times = []
        for it in range(1 + a.steps):              # one warm-up (page-locking of the Blitz storage happens here), then K steps
            Pp.fill(0.0)
            barrier(); s.synchronize()
            t0 = time.perf_counter()
            h.mg_solve_inplace()                    # H2D of Pp, mgRHS -> V-cycles -> D2H of Pp, blocking
            dt = time.perf_counter() - t0
            if it >= 1:
                e2e_times.append(allmax(dt))
        e2e_val = lups_step / float(np.mean(e2e_times)) / 1e9
        e2e = dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=int(allsum(2 * nbytes)), d2h_bytes_per_step=int(allsum(nbytes)),
                   ms_per_step=1e3 * float(np.mean(e2e_times)), api="multigrid_d3::mgSolve(plainsf&, const plainsf&) on host arrays (C++ drop-in over the C-ABI)")

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            g1, info1, t1 = reference_sample(a, 2)
            cyc = int(min(40, max(2, math.ceil(12.0 / max(t1[0] / 2.0, 1e-3)))))
            g, info, _ = reference_sample(a, cyc)
            cpu = dict(value=g, unit=UNIT, cores=info["cores"], kind=info["kind"], sample=info["sample"],
                       seconds=info["seconds_per_solve"], host_cores=info["host_cores"])
        except Exception as e:  # the baseline is a reported number, never a dependency of the GPU arm
            cpu = dict(value=None, unit=UNIT, cores=0, kind="unavailable", sample=str(e))

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms_total / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=workload_name(a, npx, npy), l2="inputs larger than L2 (level-0 arrays %.2f GB each per GPU vs 126 MB L2)" % (nbytes / 1e9),
                                cycles_per_step=int(outs[-1]["cycles"]), final_residual=float(outs[-1]["residuals"][-1]),
                                coarse_iters=[int(v) for v in outs[-1]["coarse_iters"]], lups_per_step=lups_step,
                                fine_point_vcycles_per_s=level_points[0] * outs[-1]["cycles"] * a.steps / (ms_total * 1e-3)),
                    clocks=clk, e2e=e2e, gpu_launches=int(launches), roofline=roofline, per_level=per_level, per_level_all=per_level_all,
                    time_to_tolerance=ttt, cpu_baseline=cpu, breakdown=breakdown)
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
