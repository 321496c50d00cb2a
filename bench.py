#!/usr/bin/env python
"""bench.py — V-cycle throughput of the openmg hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 3d|2d|1d]
                    [--shape 512 512 512] [--grid-levels 5] [--smoother jacobi|rbgs]
                    [--pre 1] [--post 1] [--reps 5]

Workload (default, BASELINE.json configs[3]): 3-D Poisson 512^3 (openmg's generator: diag -12,
+1 at +-1, +-NX, +-NX*NY), fp64, gridLevels=5 -> 6 grids (coarsest 16^3), V(1,1), weighted Jacobi,
u = RandomState(0).random_sample(N), b = A u, zero initial iterate.  --config 2d / 1d select
configs[2] (2-D 8192^2, 8 grids, two-colour GS) and configs[1] (1-D 2^24, 21 grids, Jacobi).
A "step" is one V-cycle.  `value` = DOF*cycles/s with b resident in HBM (CUDA events on the library
stream, exactly K cycles per repetition, best of --reps repetitions, max over ranks); `e2e` = the same
metric through the public call (Hierarchy.solve -> omg_solve) from pinned HOST buffers, host<->device
copies inside the timed region.  The line also carries the same cycle with the API's default smoother
(`rbgs`), and for N > 1 the strong-scaling number of the single-GPU problem next to the weak-scaling
`value`.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dof_cycles_per_s"
UNIT = "DOF*cycles/s"


def algorithmic_bytes_per_cycle(sizes, pre, post, with_norm=False):
    """SURVEY.md §8(d) / BASELINE.md §3: compulsory fp64 vector traffic of one fused V(pre,post)."""
    B = 0.0
    L = len(sizes) - 1
    for l in range(L):
        n, nc = sizes[l], sizes[l + 1]
        B += 24.0 * pre * n - (8.0 * n if (l >= 1 and pre >= 1) else 0.0)
        B += 16.0 * n + 8.0 * nc
        B += 8.0 * nc + (24.0 * post * n if post >= 1 else 16.0 * n)
    B += 16.0 * sizes[L]
    if with_norm:
        B += 16.0 * sizes[0]
    return B


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.idx = device_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    f = [s.strip() for s in line.split(",")]
                    if len(f) >= 9:
                        self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        mx = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, name in enumerate(names):
                if s[5 + k].lower().startswith("active"):
                    reasons.add(name)
        pw = [float(s[3]) for s in self.samples if s[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.samples), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def problem(shape, s1):
    import openmg_b200 as omg
    return omg.operators.poisson_band(shape, sparse_1d=s1)


def dist_setup(ngpus):
    """torchrun plumbing: returns (rank, world, barrier, allreduce_max)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1, (lambda: None), (lambda v: v), None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return rank, world, barrier, allmax, dist


# ---------------------------------------------------------------------------------- CPU legs

def cpu_port_rate(shape, gl, pre, post, smoother, budget_s=20.0):
    """The oracle's vectorised scipy restatement of the SAME cycle (same smoother) on the host,
    one thread, on a bounded sample of the workload (a smaller cube, same depth rule)."""
    import oracle.openmg_oracle as orc
    s = len(shape)
    sample = {3: (256, 256, 256), 2: (4096, 4096), 1: (1 << 24,)}[s]
    sample = tuple(min(a, b) for a, b in zip(sample, shape))
    A0 = orc.poisson_csr(sample, sparse_1d=(s == 1))
    N = A0.shape[0]
    u = np.random.RandomState(0).random_sample(N)
    b = A0.dot(u)
    t0 = time.perf_counter()
    R = orc.restrictionList(sample, gl - 1, 8)
    A = orc.coeffecientList(A0, R)
    setup_s = time.perf_counter() - t0
    params = {'coarsestLevel': len(R), 'preIterations': pre, 'postIterations': post, 'verbose': False}
    smooth = orc.make_smoother(smoother, sample, 0.8)
    x = None
    cycles = 0
    t0 = time.perf_counter()
    while True:
        x, info = orc.mgCycle(A, b, 0, R, params, initial=x, smooth=smooth)
        cycles += 1
        el = time.perf_counter() - t0
        if el > budget_s or cycles >= 8:
            break
    rate = N * cycles / el
    return {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle (scipy CSR, vectorised %s) V(%d,%d) on %s, %d grids, %d cycles in %.1f s "
                      "(+%.1f s setup), host has %d cores"
                      % (smoother, pre, post, "x".join(map(str, sample)), len(A), cycles, el, setup_s,
                         os.cpu_count())}


def workload_config(shape, gl, pre, post):
    """`config` of the JSON line: identical in both arms (the reference arm times a bounded sample of it)."""
    return {"workload": "%d-D Poisson %s fp64 (openmg generator), gridLevels=%d, V(%d,%d), b=A*u u~U[0,1), zero "
                        "initial iterate" % (len(shape), "x".join(map(str, shape)), gl, pre, post)}


def reference_arm(args):
    """--impl reference: the reference's own algorithm as-is (lexicographic Gauss-Seidel in a
    Python loop over CSR rows, openmg/solvers.py:34-75; SuperLU coarse solve every cycle) through
    the oracle port (the reference is Python 2 and does not exist on the GPU box), each step one
    V-cycle on a bounded sample of the workload.  Single thread by construction: the reference's
    hot loop is a pure-Python `for i in range(N)`."""
    import oracle.openmg_oracle as orc
    shape = tuple(args.shape)
    if args.gpus > 1 and args.scaling == "weak" and shape == (512, 512, 512) and args.gpus in WEAK_SHAPES:
        shape, args.grid_levels = WEAK_SHAPES[args.gpus]          # the same workload as our arm at this N
    s = len(shape)
    sample = {3: (24, 24, 24), 2: (128, 128), 1: (1 << 14,)}[s]
    sample = tuple(min(a, b) for a, b in zip(sample, shape))
    A0 = orc.poisson_csr(sample, sparse_1d=(s == 1))
    N = A0.shape[0]
    u = np.random.RandomState(0).random_sample(N)
    b = A0.dot(u)
    R = orc.restrictionList(sample, args.grid_levels - 1, 8)
    A = orc.coeffecientList(A0, R)
    params = {'coarsestLevel': len(R), 'preIterations': args.pre, 'postIterations': args.post, 'verbose': False}
    smooth = orc.make_smoother('gs', sample, fast=False)          # the literal Python row loop
    x = None
    for _ in range(args.warmup):
        x, _i = orc.mgCycle(A, b, 0, R, params, initial=x, smooth=smooth)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        x, _i = orc.mgCycle(A, b, 0, R, params, initial=x, smooth=smooth)
    el = time.perf_counter() - t0
    rate = N * args.steps / el
    port = cpu_port_rate(shape, args.grid_levels, args.pre, args.post, args.smoother, budget_s=10.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "vcycles_per_s": args.steps / el,
        "config": workload_config(shape, args.grid_levels, args.pre, args.post),
        "smoother": "lexicographic GS (reference as-is)",
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "reference algorithm as-is (pure-Python lexicographic GS row loop + spsolve) on "
                                   "%s, %d grids, %d timed + %d warm-up cycles; single thread by construction; host "
                                   "has %d cores" % ("x".join(map(str, sample)), len(A), args.steps, args.warmup,
                                                     os.cpu_count())},
        "cpu_port_same_smoother": port,
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------- our arm

WEAK_SHAPES = {   # per-GPU work fixed at 512^3 points; (a, b, a) shapes keep the reference's offsets consistent
    1: ((512, 512, 512), 5),
    2: ((512, 1024, 512), 6),
    4: ((512, 2048, 512), 6),
    8: ((1024, 1024, 1024), 6),      # BASELINE.json configs[4]
}


def timed_cycles(h, args, smoother, barrier, allmax, reps):
    """W warm-up cycles, then `reps` repetitions of exactly K device-resident cycles (CUDA events on the library
    stream inside omg_bench_cycles, barrier + synchronize on both sides, max over ranks).  Returns the per-repetition
    times (ms per K cycles) and the launch count of one repetition."""
    h.bench_cycles(max(args.warmup, 3), args.pre, args.post, smoother, 0.8)
    out, launches = [], 0
    for _ in range(reps):
        barrier()
        ms, launches = h.bench_cycles(args.steps, args.pre, args.post, smoother, 0.8)
        barrier()
        out.append(allmax(ms))            # identical on every rank from here on (collectives must match)
    return out, launches


def ours(args):
    rank, world, barrier, allmax, dist = dist_setup(args.gpus)
    import openmg_b200 as omg
    from openmg_b200 import _lib
    from openmg_b200.hierarchy import Hierarchy

    shape, gl = tuple(args.shape), args.grid_levels
    base_shape, base_gl = shape, gl
    scaling = "weak"
    if world > 1:
        from openmg_b200 import dist as omg_dist
        omg_dist.init_from_torch(dist)
        if args.scaling == "weak" and shape == (512, 512, 512) and world in WEAK_SHAPES:
            shape, gl = WEAK_SHAPES[world]
        else:
            scaling = "strong"
    s1 = len(shape) == 1
    dev = _lib.device_info()
    A = problem(shape, s1)
    N = A.n
    t0 = time.perf_counter()
    h = Hierarchy(A, shape, gl - 1, 8, flags=_lib.FLAG_FORCE_CSR if args.force_csr else 0)
    setup_wall = time.perf_counter() - t0
    nlev = h.nlevels
    sizes = [h.level_info(l)["n"] for l in range(nlev)]
    kinds = [h.level_info(l)["kind"] for l in range(nlev)]
    row0, nloc, slab = h.local_range(0)
    slabs = [h.local_range(l)[2] for l in range(nlev)]

    # synthetic input: u uniform[0,1), b = A u; b computed on the device from host u.
    # One GPU: u = RandomState(0).random_sample(N).  N GPUs: each rank draws its own rows (seed = rank).
    u = np.random.RandomState(rank if world > 1 else 0).random_sample(nloc)
    b_host = _lib.pinned_empty(nloc)
    b_host[:] = h.matvec_local(u, 0)
    x_host = _lib.pinned_empty(nloc)
    del u
    h.set_rhs_local(b_host)

    # ---- device-resident timing
    with ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) as cs:
        rep_ms, launches = timed_cycles(h, args, args.smoother, barrier, allmax, args.reps)
        ms = min(rep_ms)
        # per-kernel shares (CUDA events around every launch of 3 directly launched cycles), taken right behind the timed
        # repetitions: under the power cap a B200 slows by several per cent within the first second of sustained load
        # (see `repetitions_ms_per_step`), so a profile taken after the filler below would describe a hotter GPU than
        # the one `value` was measured on
        prof = h.profile_cycle(3, args.pre, args.post, args.smoother, 0.8)
        if sum(rep_ms) < 1500:   # keep the GPU under the same load a little longer so the sampler sees it
            h.bench_cycles(max(args.steps, int(1500 / max(ms / args.steps, 1e-3))), args.pre, args.post,
                           args.smoother, 0.8)
    clocks = cs.summary()
    final_norm = h.current_norm()
    value = N * args.steps / (ms * 1e-3)
    peak, peak_src = peaks()
    Bcyc = algorithmic_bytes_per_cycle(sizes, args.pre, args.post)
    if args.force_csr:
        # general-matrix path: every application of a level operator also streams the matrix, 12 B per stored
        # entry (fp64 value + int32 column) + 12 B per row (row length, a_ii)
        nnzs = [h.level_info(l)["nnzA"] for l in range(nlev)]
        for l in range(nlev - 1):
            Bcyc += (args.pre + 1 + args.post) * (12.0 * nnzs[l] + 12.0 * sizes[l])

    def cycle_roofline(ms_k, moved=None):
        # Bcyc is SURVEY section 8(d)'s formula (every sweep 24 n, the residual+restriction 16 n + 8 n_c), kept as the
        # denominator so the fraction stays comparable with BASELINE.md and round 1.  The kernels that fuse the last
        # pre-smoothing sweep with the residual (k_jr3) or start from the zero iterate without reading x move LESS than
        # that; `fused_bytes_per_cycle` is what the launches of this cycle have to move at the least (sum of the
        # per-launch algorithmic bytes of the `kernels` list), and `fused_frac_of_measured_peak` the fraction of the copy
        # peak the cycle reaches counting only those.
        ach = Bcyc / (ms_k / args.steps * 1e-3) / 1e9 / world
        out = {"algorithmic_bytes_per_cycle": Bcyc, "achieved_per_gpu": ach, "unit": "GB/s",
               "frac_of_measured_peak": ach / peak, "frac_of_8TBs_nominal": ach / 8000.0}
        if moved:
            fach = moved / (ms_k / args.steps * 1e-3) / 1e9
            out.update({"fused_bytes_per_cycle_per_gpu": moved, "fused_achieved_per_gpu": fach,
                        "fused_frac_of_measured_peak": fach / peak})
        return out

    # ---- the roofline of the dominant kernel
    tot = sum(p["ms"] * p["launches"] / 3.0 for p in prof)
    dom = max(prof, key=lambda p: p["ms"] * p["launches"] if p["bytes"] > 0 else 0.0)
    ach = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "%s@L%d" % (dom["name"], dom["level"]), "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "share_of_cycle": dom["ms"] * dom["launches"] / 3.0 / tot if tot > 0 else None,
                "algorithmic_bytes_per_launch": dom["bytes"]}
    # DRAM bytes per launch of that kernel: not measurable inside a timed run (it needs ncu's dram__bytes counters);
    # taken from the committed `ncu --set full` capture of the same command when its shape matches, else null
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            ent = tj.get("x".join(map(str, shape)), {}).get("%s@L%d" % (dom["name"], dom["level"]))
            if ent is not None:
                roofline["traffic"] = ent
                roofline["traffic_source"] = tj.get("source")
        except Exception:  # noqa: BLE001
            pass

    # ---- the same cycle with the other smoother ('rbgs' is the default of the Python API)
    other = None
    if not args.no_extra:
        osm = "rbgs" if args.smoother == "jacobi" else "jacobi"
        o_ms, o_launches = timed_cycles(h, args, osm, barrier, allmax, max(2, args.reps // 2))
        o_best = min(o_ms)
        other = {"smoother": osm, "value": N * args.steps / (o_best * 1e-3), "unit": UNIT,
                 "ms_per_step": o_best / args.steps, "vcycles_per_s": args.steps / (o_best * 1e-3),
                 "gpu_launches": int(o_launches), "cycle_roofline": cycle_roofline(o_best)}
        h.set_rhs_local(b_host)

    # ---- end to end through the public call with HOST buffers (pinned), copies inside the timed region
    cyc_call = args.e2e_cycles
    h.solve_local(b_host, x_host, args.pre, args.post, args.smoother, 0.8, 1, 0.0)       # warm
    barrier()
    ncalls = max(1, args.steps // cyc_call)
    t0 = time.perf_counter()
    for _ in range(ncalls):
        done, norm = h.solve_local(b_host, x_host, args.pre, args.post, args.smoother, 0.8, cyc_call, 0.0)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e = {"value": N * cyc_call * ncalls / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": 8.0 * N / cyc_call, "d2h_bytes_per_step": 8.0 * N / cyc_call,
           "cycles_per_call": cyc_call, "calls": ncalls, "ms_per_call": 1e3 * e2e_s / ncalls,
           "call": "openmg_b200.Hierarchy.solve -> omg_solve (pinned host b in, host x out, final residual norm "
                   "read; a step is one V-cycle, bytes are per cycle summed over ranks)",
           "final_norm": norm}
    setup_times = h.setup_times()
    h.close()
    del h

    # ---- N > 1, weak-scaled default workload: the strong-scaling number of the single-GPU problem as well
    strong = None
    if world > 1 and scaling == "weak" and not args.no_extra:
        A1 = problem(base_shape, len(base_shape) == 1)
        h1 = Hierarchy(A1, base_shape, base_gl - 1, 8)
        r0, nl, _sl = h1.local_range(0)
        u = np.random.RandomState(rank).random_sample(nl)
        h1.set_rhs_local(h1.matvec_local(u, 0))
        del u
        s_ms, s_launches = timed_cycles(h1, args, args.smoother, barrier, allmax, max(2, args.reps // 2))
        s_best = min(s_ms)
        strong = {"workload": workload_config(base_shape, base_gl, args.pre, args.post)["workload"],
                  "value": A1.n * args.steps / (s_best * 1e-3), "unit": UNIT, "ms_per_step": s_best / args.steps,
                  "vcycles_per_s": args.steps / (s_best * 1e-3), "scaling": "strong",
                  "level_is_slab": [h1.local_range(l)[2] for l in range(h1.nlevels)]}
        h1.close()

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = cpu_port_rate(base_shape, base_gl, args.pre, args.post, args.smoother,
                        budget_s=20.0 if world == 1 else 8.0)
    cfg = workload_config(shape, gl, args.pre, args.post)
    if args.force_csr:
        cfg["workload"] += "; general-matrix path (every level an explicit CSR operator in sliced-ELLPACK form)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "vcycles_per_s": args.steps / (ms * 1e-3),
        "config": cfg,
        "smoother": "%s (omega=0.8)" % args.smoother if args.smoother == "jacobi" else "two-colour GS (rbgs)",
        "details": {"grids": nlev, "level_rows": sizes, "level_kinds": kinds, "level_is_slab": slabs,
                    "l2_policy": "inputs larger than L2 (3 x %.2f GB level-0 vectors per GPU)" % (8e-9 * nloc),
                    "parallelism": "slab%d" % world, "device": dev["name"],
                    "repetitions_ms_per_step": [m / args.steps for m in rep_ms], "timing": "best of %d repetitions "
                    "of exactly %d cycles" % (len(rep_ms), args.steps)},
        "roofline": roofline,
        "cycle_roofline": cycle_roofline(ms, sum(p["bytes"] * p["launches"] / 3.0 for p in prof)),
        "kernels": [{"kernel": "%s@L%d" % (p["name"], p["level"]), "launches_per_cycle": p["launches"] / 3.0,
                     "ms": p["ms"], "GBs": p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else None}
                    for p in prof],
        "other_smoother": other, "strong_scaling_same_problem": strong,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "setup": dict(setup_times, wall_s=setup_wall), "final_norm_after_timed_cycles": final_norm,
    }
    print(json.dumps(line))


CONFIGS = {     # BASELINE.json configs[1..3]: shape, gridLevels, smoother
    "3d": ((512, 512, 512), 5, "jacobi"),
    "2d": ((8192, 8192), 7, "rbgs"),
    "1d": ((1 << 24,), 20, "jacobi"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="3d", choices=sorted(CONFIGS),
                    help="BASELINE.json workload: 3d = configs[3] (default), 2d = configs[2], 1d = configs[1]")
    ap.add_argument("--shape", type=int, nargs="+", default=None)
    ap.add_argument("--grid-levels", type=int, default=None)
    ap.add_argument("--smoother", default=None, choices=["jacobi", "rbgs"])
    ap.add_argument("--reps", type=int, default=5, help="repetitions of the K timed cycles (best is reported)")
    ap.add_argument("--no-extra", action="store_true", help="skip the other-smoother and strong-scaling legs")
    ap.add_argument("--force-csr", action="store_true",
                    help="run the general-matrix path: every level as explicit CSR (no band fast path)")
    ap.add_argument("--pre", type=int, default=1)
    ap.add_argument("--post", type=int, default=1)
    ap.add_argument("--e2e-cycles", type=int, default=10)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = 512^3 points per GPU (default), strong = the same --shape on N GPUs")
    args = ap.parse_args()
    cshape, cgl, csm = CONFIGS[args.config]
    args.shape = list(cshape) if args.shape is None else args.shape
    args.grid_levels = cgl if args.grid_levels is None else args.grid_levels
    args.smoother = csm if args.smoother is None else args.smoother
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        reference_arm(args)
        return 0
    ours(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
