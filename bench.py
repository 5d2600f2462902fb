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
lib/poisson compiled unmodified with its -Ofast flags, run with one thread-rank per core.  Those legs run in a CHILD
process (`--_cpu-leg`): the reference's fatal paths call exit(0) from inside a thread-rank (parser.cc:250-352), which no
try/except can contain, and the GPU line must not depend on them.

`config` holds only what the command line determines, so that both arms print the same object; everything measured in
the run (cycles, residual, LUPs) is under `run`.
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
BYTES_PER_LUP = {"smooth": 24.0, "residual": 24.0, "coarsen": 9.0, "prolong": 17.0, "error": 16.0,
                 "residual+coarsen": 17.0}     # the fused pass of the largest levels: read p, read rhs, write 1/8 coarse rhs
GS_KERNEL = "k_gs_skew<3> lexicographic Gauss-Seidel"


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
    ap.add_argument("--cpu-index", type=int, default=0, help="grid index of the bounded CPU sample (0: the largest of 9 / 8 that fits the time budget)")
    ap.add_argument("--no-proj", action="store_true", help="skip the device-resident projection-step leg (divergence -> mgSolve -> correction, SURVEY 8f-1/8f-2)")
    ap.add_argument("--proj-multi", action="store_true", help="run the projection-step leg on N > 1 ranks as well")
    ap.add_argument("--no-gs", action="store_true", help="skip the Gauss-Seidel leg (the reference YAMLs' smoother) of the default Jacobi run")
    ap.add_argument("--_cpu-leg", dest="cpu_leg", default=None, help=argparse.SUPPRESS)
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
def cpu_layout(cores, index, depth):
    """Largest power-of-two npX x npY <= cores that the reference accepts: local intervals per split direction must stay
    >= 2^(vcDepth+1) (parser.cc:305-337, otherwise checkData prints and calls exit(0)), i.e. ranks per axis <= 2^(index-depth-1)."""
    per_axis = max(1, 1 << max(0, index - depth - 1))
    npx = npy = 1
    while True:
        if npx <= npy and npx * 2 <= per_axis and npx * 2 * npy <= min(cores, 64):
            npx *= 2
        elif npy * 2 <= per_axis and npx * npy * 2 <= min(cores, 64):
            npy *= 2
        elif npx * 2 <= per_axis and npx * 2 * npy <= min(cores, 64):
            npx *= 2
        else:
            break
    if npy > npx:
        npx, npy = npy, npx
    return npx, npy


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def reference_solve_times(a, index, cycles, npx, npy, repeats=1, warm=0):
    """Times poisson::mgSolve of the reference (oracle/_ref, else the oracle port) on a `index` grid.  Runs in the child."""
    from oracle import sarasref
    kind = "reference" if sarasref.available(3, 0, fast=True) else "port"
    depth = index - 3
    if kind != "reference":
        npx, npy = 1, 1
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
    out = None
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
    sample = "%d^3 grid (Index %d), %d V(2,2) cycles per solve, depth %d, %s smoother, %dx%d thread-ranks, %s" % (
        n[2], index, out["cycles"], depth, a.smoother, npx, npy,
        "reference lib/poisson compiled -Ofast (oracle/_ref)" if kind == "reference" else "oracle/mg_oracle.c -O3")
    return dict(kind=kind, cores=nr, sample=sample, times=times, lups_per_solve=lups, cycles=int(out["cycles"]))


def cpu_leg_child(a):
    """Child process: spec = JSON {mode, budget_s, steps, warmup}.  Prints one JSON object."""
    spec = json.loads(a.cpu_leg)
    cores = host_cores()
    budget = float(spec.get("budget_s", 25.0))
    steps, warm = int(spec.get("steps", 1)), int(spec.get("warmup", 0))
    res = dict(host_cores=cores)
    # probe: one cycle at 257^3 on every core the reference can use
    npx8, npy8 = cpu_layout(cores, 8, 5)
    probe = reference_solve_times(a, 8, 1, npx8, npy8)
    t_cycle8 = probe["times"][0]
    res["probe"] = dict(sample=probe["sample"], seconds=t_cycle8)
    # choose the sample: 513^3 (the largest grid the CPU arm can afford) if `steps + warmup` solves of >= 2 cycles fit the budget
    index = a.cpu_index
    if index == 0:
        npx9, npy9 = cpu_layout(cores, 9, 6)
        t_cycle9 = t_cycle8 * 8.0 * (npx8 * npy8) / (npx9 * npy9)
        index = 9 if (steps + warm) * 2 * t_cycle9 <= budget else 8
    depth = index - 3
    npx, npy = cpu_layout(cores, index, depth)
    t_cycle = t_cycle8 * (8.0 ** (index - 8)) * (npx8 * npy8) / (npx * npy)
    cycles = int(max(1, min(a.cycles, budget / max((steps + warm) * t_cycle, 1e-6))))
    r = reference_solve_times(a, index, cycles, npx, npy, repeats=steps, warm=warm)
    sec = float(np.mean(r["times"]))
    res.update(kind=r["kind"], cores=r["cores"], sample=r["sample"], seconds_per_solve=sec, times=r["times"],
               value=r["lups_per_solve"] / sec / 1e9, lups_per_solve=r["lups_per_solve"], layout=[npx, npy], index=index)
    if spec.get("single_core", False):
        # np = 1 on one core: 2 cycles at 257^3 (about 3 s)
        r1 = reference_solve_times(a, 8, 2, 1, 1)
        res["single_core"] = dict(value=r1["lups_per_solve"] / r1["times"][0] / 1e9, unit=UNIT, cores=1, sample=r1["sample"], seconds=r1["times"][0])
    print("CPU_LEG_JSON " + json.dumps(res), flush=True)


def run_cpu_leg(a, spec, timeout):
    """Runs the CPU leg in a child process and returns its result dict (or a dict with kind == 'unavailable')."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--_cpu-leg", json.dumps(spec),
           "--index", str(a.index), "--cycles", str(a.cycles), "--smoother", a.smoother, "--mesh", a.mesh,
           "--tolerance", repr(a.tolerance), "--cpu-index", str(a.cpu_index)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return dict(kind="unavailable", why="CPU leg exceeded %d s" % timeout)
    for line in r.stdout.splitlines():
        if line.startswith("CPU_LEG_JSON "):
            return json.loads(line[len("CPU_LEG_JSON "):])
    return dict(kind="unavailable", why="CPU leg ended with rc %d and no result: %s" % (r.returncode, (r.stderr or r.stdout)[-300:].replace("\n", " | ")))


def bench_config(a, npx, npy, level0_bytes=None):
    """what the command line determines -- identical in both arms"""
    n = (1 << a.index) + 1
    if level0_bytes is None:
        level0_bytes = ((n - 1) // npx + 3) * ((n - 1) // npy + 3) * (n + 2) * 8
    return dict(workload=workload_name(a, npx, npy),
                l2="inputs larger than L2 (level-0 arrays %.2f GB each per GPU vs 126 MB L2)" % (level0_bytes / 1e9),
                cycles_per_solve_max=a.cycles, tolerance=a.tolerance)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    npx, npy = LAYOUTS.get(a.gpus, (1, 1))
    res = run_cpu_leg(a, dict(steps=a.steps, warmup=a.warmup, budget_s=120.0), timeout=900)
    if res.get("kind") == "unavailable":
        # last resort: one rank at 257^3 (always a valid layout)
        a.cpu_index = 8
        res = run_cpu_leg(a, dict(steps=a.steps, warmup=a.warmup, budget_s=60.0), timeout=600)
    if res.get("kind") == "unavailable":
        print(json.dumps(dict(impl="reference", unavailable=res.get("why", "CPU leg failed"))), flush=True)
        return
    glups = res["value"]
    line = dict(metric=METRIC, value=glups, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=1e3 * res["seconds_per_solve"],
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=bench_config(a, npx, npy),
                cpu_baseline=dict(value=glups, unit=UNIT, cores=res["cores"], kind=res["kind"], sample=res["sample"], host_cores=res["host_cores"]),
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

    # per-operator device times of the fine levels, measured live with CUDA events over the timed region (the event pairs of
    # levels <= 1 are recorded INSIDE the timed region: ~60 records per cycle, no synchronisation)
    peak, peak_src = measured_peak_hbm()

    def fine_level_table():
        rows = []
        for lvl in range(min(2, depth + 1)):
            for kind, name in ((capi.K_SMOOTH, "smooth"), (capi.K_RESIDUAL, "residual"), (capi.K_COARSEN, "coarsen"),
                               (capi.K_FUSED_RC, "residual+coarsen"), (capi.K_PROLONG, "prolong"), (capi.K_ERROR, "error")):
                ms, n = s.profile_query(kind, lvl)
                if n == 0:
                    continue
                if name == "residual+coarsen":
                    n //= 2                 # two event pairs per fused pass (shell residual + fused kernel, shell restriction)
                pts = level_points_local[lvl]
                gbs = BYTES_PER_LUP[name] * pts / (ms / n * 1e-3) / 1e9
                rows.append(dict(level=lvl, op=name, launches=int(n), avg_ms=ms / n, gbs=gbs, frac=gbs / peak))
        return rows

    def roofline_of(rows, smoother):
        dom = next((e for e in rows if e["level"] == 0 and e["op"] == "smooth"), None)
        if not dom:
            return None
        traffic = None      # DRAM bytes per launch of this kernel from the committed ncu capture (same grid, same layout only)
        if a.gpus == 1 and a.index == 10 and a.mesh == "uniform":
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_1025_n1.json")))[smoother]["bytes"]
            except Exception:
                traffic = None
        return dict(bound="hbm", achieved=dom["gbs"], peak=peak, unit="GB/s", frac=dom["gbs"] / peak, traffic=traffic,
                    traffic_source="profiles/traffic_1025_n1.json (ncu --set full capture of this kernel, 1 GPU, 1025^3 uniform); null for other layouts",
                    kernel="level-0 smoother sweep (%s)" % (GS_KERNEL if smoother == "gs" else "k_stencil_tma<0> Jacobi, TMA-staged x-march"),
                    algorithmic_bytes_per_launch=BYTES_PER_LUP["smooth"] * level_points_local[0], avg_launch_ms=dom["avg_ms"],
                    peak_source=peak_src)

    per_level = fine_level_table()
    roofline = roofline_of(per_level, a.smoother)

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
                           (capi.K_FUSED_RC, "residual+coarsen"), (capi.K_PROLONG, "prolong"), (capi.K_ERROR, "error"), (capi.K_HALO, "halo"),
                           (capi.K_BC, "bc"), (capi.K_COARSE_SOLVE, "coarse_solve")):
            ms, n = s.profile_query(kind, lvl)
            if n == 0:
                continue
            if name == "residual+coarsen":
                n //= 2
            e = dict(level=lvl, op=name, launches=int(n), avg_us=1e3 * ms / n)
            if name in BYTES_PER_LUP:
                gbs = BYTES_PER_LUP[name] * level_points_local[lvl] / (ms / n * 1e-3) / 1e9
                e.update(gbs=gbs, frac=gbs / peak)
            per_level_all.append(e)
    s.set_profile_max_level(1)

    # projection step with every field resident: V.divergence -> mgRHS *= 1/dt -> mgSolve -> Pp.syncData ->
    # P += Pp -> V -= dt grad(Pp) per step, nothing but the residual history crossing PCIe.  Reported beside `e2e`: it is what
    # the hot path costs once hydro_d3's fields live on the device (uniform meshes only here: the metrics are 1/L).
    def projection_leg():
        n0 = dims[0]
        shapes = s.proj_shapes()
        L = 1.0
        stag = [np.full(n0[d] + 2, 1.0 / L) for d in range(3)]
        coll = [np.full(n0[d] + 1, 1.0 / L) for d in range(3)]
        spacing = [1.0 / (1 << a.index)] * 3
        s.proj_setup(stag, coll, spacing)
        src = mgRHS.reshape(-1)             # random in [-1, 1): reused as the three velocity components
        for which in (capi.PROJ_VX, capi.PROJ_VY, capi.PROJ_VZ):
            cnt = int(np.prod(shapes[which]))
            s.proj_upload(which, src[:cnt].reshape(shapes[which]))
        dt = 0.01
        s.proj_step(dt)                     # warm-up
        barrier(); s.synchronize()
        s.event_record(2)
        s.proj_divergence(1.0 / dt)
        s.event_record(3)
        s.proj_correct(dt)
        s.event_record(4)
        ms_div, ms_cor = allmax(s.event_elapsed_ms(2, 3)), allmax(s.event_elapsed_ms(3, 4))
        steps = max(2, min(a.steps, 3))
        barrier(); s.synchronize()
        s.event_record(5)
        outs_p = [s.proj_step(dt) for _ in range(steps)]
        s.event_record(6)
        ms = allmax(s.event_elapsed_ms(5, 6)) / steps
        s.synchronize(); barrier()
        lups = sum(lups_per_solve(level_points, pre, post, o["coarse_iters"]) for o in outs_p) / steps
        pts = level_points_local[0]
        # implicit solve of Vx (hydro_d3::solveVx with the lid-driven-cavity boundary objects), one rank: Jacobi sweeps
        # of 40 B per point (update 24, max-norm of the implicit equation 16) plus the boundary planes, one scalar read back per sweep
        implicit = None
        if world == 1:
            try:
                bt, bv = capi.velocity_bcs(0, 0, 0, 1)
                idt, iRe = 1.0e-5, 1000.0           # dt / (2 Re h^2) = 5e-3 at h = 2^-10: converges in a few sweeps
                s.impl_setup(idt, iRe, 1.0e-6, spacing, [np.full(n0[d] + 1, 1.0 / (L * L)) for d in range(3)],
                             [np.full(n0[d] + 2, 1.0 / (L * L)) for d in range(3)], bt, bv, 1000)
                cnt = int(np.prod(shapes[capi.PROJ_VX]))
                s.impl_upload_rhs(0, src[:cnt].reshape(shapes[capi.PROJ_VX]))
                s.synchronize()
                s.event_record(2)
                sweeps, err = s.impl_solve(0)
                s.event_record(3)
                ms_i = s.event_elapsed_ms(2, 3)
                implicit = dict(component="Vx", sweeps=int(sweeps), max_error=float(err), ms=ms_i, ms_per_sweep=ms_i / sweeps,
                                gbs=40.0 * cnt / (ms_i / sweeps * 1e-3) / 1e9, dt=idt, Re=iRe, cn_tolerance=1.0e-6,
                                bytes_per_point="40 per sweep: Jacobi update 24 (read V, rhs, write guess) + residual max-norm 16",
                                api="sarasmg_impl_solve (hydro_d3.cc:370-423 + hydro.cc:386-401)")
            except Exception as e:
                implicit = dict(error=str(e)[:300])
        return dict(implicit_solve=implicit, ms_per_step=ms, value=lups / (ms * 1e-3) / 1e9, unit=UNIT, cycles_per_step=[int(o["cycles"]) for o in outs_p],
                    divergence_ms=ms_div, correction_ms=ms_cor,
                    divergence_gbs=32.0 * pts / (ms_div * 1e-3) / 1e9, correction_gbs=80.0 * pts / (ms_cor * 1e-3) / 1e9,
                    bytes_per_point="divergence 32 (read Vx, Vy, Vz, write mgRHS); correction 80 (copy Pp 16, P += Pp 24, three velocity updates 48 - 8 shared Pp read)",
                    pcie_bytes_per_step=0,
                    api="sarasmg_proj_step: vfield::divergence, mgRHS *= 1/dt, poisson::mgSolve, Pp.syncData, P += Pp, V -= dt grad(Pp) (hydro_d3.cc:326-362), all fields device-resident")

    # time to tolerance: the same hierarchy run until the reference's stagnation test |r_prev - r| < mgTolerance stops it
    def time_to_tolerance():
        s.set_vc_count(60)
        s.reset_counters()
        barrier(); s.synchronize()
        s.event_record(2)
        out_t = s.solve_resident()
        s.event_record(3)
        ms_t = allmax(s.event_elapsed_ms(2, 3))
        s.set_vc_count(a.cycles)
        return dict(ms=ms_t, cycles=int(out_t["cycles"]), tolerance=a.tolerance, final_residual=float(out_t["residuals"][-1]),
                    stagnated=bool(out_t["cycles"] < 60), rule="poisson.cc:145 stagnation test, at most 60 cycles")

    ttt = None
    if not a.no_ttt:
        ttt = {a.smoother: time_to_tolerance()}

    # the reference YAMLs' smoother on the same resident hierarchy: lexicographic Gauss-Seidel (Smoothing Method: 1 in input/,
    # tests/mgTest, tests/ldcTest).  Same problem, same timing rules; reported beside the Jacobi headline.
    gs = None
    if a.smoother == "jacobi" and not a.no_gs:
        s.set_smoother(1)
        for _ in range(2):
            out_g = s.solve_resident()
        s.reset_counters()
        barrier(); s.synchronize()
        gsteps = max(2, min(a.steps, 5))
        s.event_record(4)
        for _ in range(gsteps):
            out_g = s.solve_resident()
        s.event_record(5)
        ms_g = allmax(s.event_elapsed_ms(4, 5))
        lups_g = lups_per_solve(level_points, pre, post, out_g["coarse_iters"])
        rows_g = fine_level_table()
        gs = dict(value=lups_g * gsteps / (ms_g * 1e-3) / 1e9, unit=UNIT, steps=gsteps, ms_per_step=ms_g / gsteps,
                  cycles_per_step=int(out_g["cycles"]), final_residual=float(out_g["residuals"][-1]),
                  roofline=roofline_of(rows_g, "gs"), per_level=rows_g)
        if not a.no_ttt:
            ttt["gs"] = time_to_tolerance()
        s.set_smoother(0)

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

    # ---- CPU baseline beside it (rank 0, N = 1 only), in a child process --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = run_cpu_leg(a, dict(steps=1, warmup=0, budget_s=20.0, single_core=True), timeout=600)
        if r.get("kind") == "unavailable":
            cpu = dict(value=None, unit=UNIT, cores=0, kind="unavailable", sample=r.get("why", ""))
        else:
            cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"],
                       seconds=r["seconds_per_solve"], host_cores=r["host_cores"], single_core=r.get("single_core"))

    # last: the projection leg overwrites the resident inputs of solve_resident
    projection = None
    if not a.no_proj and a.mesh == "uniform" and (world == 1 or a.proj_multi):   # N > 1 only on request: an exception on one rank would leave the others in a collective
        try:
            projection = projection_leg()
        except Exception as e:            # extra information only: never take the bench line down
            projection = dict(error=str(e)[:300])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms_total / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                    data="synthetic", config=bench_config(a, npx, npy, nbytes),
                    run=dict(cycles_per_step=int(outs[-1]["cycles"]), final_residual=float(outs[-1]["residuals"][-1]),
                             coarse_iters=[int(v) for v in outs[-1]["coarse_iters"]], lups_per_step=lups_step,
                             fine_point_vcycles_per_s=level_points[0] * outs[-1]["cycles"] * a.steps / (ms_total * 1e-3)),
                    clocks=clk, e2e=e2e, gpu_launches=int(launches), roofline=roofline, per_level=per_level, per_level_all=per_level_all,
                    gauss_seidel=gs, time_to_tolerance=ttt, cpu_baseline=cpu, breakdown=breakdown, projection_step=projection)
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.cpu_leg:
        cpu_leg_child(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
