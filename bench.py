#!/usr/bin/env python
"""Benchmark of the fast-poisson-solver hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-strong] [--no-cpu]
    torchrun ... bench.py --gpus N ...            (one rank per GPU, NCCL)

One step = one pass of the hot path over one batch of synthetic input:
    Solver.precompute(x_pde, y_pde, x_bc, y_bc)  +  Solver.solve(f, u_bc)
on BASELINE.json configs[1]: 1024x1024 scattered (scrambled Sobol) collocation points in the unit
square, 4100 boundary-ring points, Perlin source, constant Dirichlet value.  With N > 1 GPUs every rank
owns its own 2^20 points (weak scaling); the Gram / boundary blocks are all-reduced once per
precompute and the projected right-hand side once per solve.

Prints ONE JSON line (rank 0):
  value          points/s, inputs resident in HBM, per-kernel instrumentation OFF
  e2e            the same step fed from pinned host buffers with the result copied back
  roofline       the dominant kernel (hidden 700x700 layer) timed with CUDA events on its own stream in a second,
                 instrumented pass over the same steps; peak = MEASURED_PEAKS.json
  int8_peak      dense int8 tensor-core rate measured in this run (torch._int_mm 8192^3, burst + sustained): the
                 denominator for the exact Gram / projection / reconstruction kernels
  batched_solve  BASELINE configs[2] (4096 right-hand sides on a 512x512 grid) per rank
  strong         BASELINE configs[2] with the right-hand sides sharded over the N GPUs, configs[3] (4.19 M points,
                 L-shape, point-sharded) and configs[4] (16.7 M points + 4096 solves, FD accuracy on 2048^2) at N GPUs
  multi_gpu_parity (N > 1) a point-sharded precompute + solve against the same problem on rank 0 alone
  cpu_baseline   the oracle's algorithm-faithful port of the reference (2 800 autograd sweeps) on this box's host cores
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 1024                      # configs[1]: 1024 x 1024 points
N_POINTS = GRID * GRID
FLOP_PER_POINT_HIDDEN = 2 * 4 * 700 * 700          # one hidden layer, 4 streams
FLOP_PER_POINT_FEATURES = 2 * 4 * (400 * 700 + 5 * 700 * 700)  # 21.84 MFLOP/point
WORKLOAD = "configs[1]: 1024x1024 scattered (Sobol) points, Perlin source, precompute + single solve"


def bench_config(world, n_pts=N_POINTS):
    """The `config` object of the JSON line - identical for both arms (--impl ours / reference)."""
    return {"workload": WORKLOAD, "points_per_gpu": n_pts, "boundary_points": 4 * GRID + 4,
            "weights": "default init, seed 0 (the pretrained final.pt is absent from the reference checkout)",
            "lambda": 2 ** -12, "parallelism": f"points sharded x{world}, 1 all-reduce per precompute + 1 per solve",
            "l2": "inputs (H, DH: 2.9 GB each) larger than L2; no flush needed"}


# ---------------------------------------------------------------------------------------------
# synthetic inputs (host side: the scrambled Sobol sequence of data.py:200-205 comes from scipy; everything else is
# generated on the device, fast_poisson_solver_b200/sources.py)
# ---------------------------------------------------------------------------------------------
def sobol_points(n, seed):
    from scipy.stats import qmc

    eng = qmc.Sobol(d=2, scramble=True, seed=seed)  # as data.py:200-205
    s = eng.random(n)
    return np.ascontiguousarray(s[:, 0]), np.ascontiguousarray(s[:, 1])


def ring(G):
    g = np.linspace(0.0, 1.0, G + 2)
    yg = g[1:-1]
    return (np.concatenate([g, g, np.zeros_like(yg), np.ones_like(yg)]),
            np.concatenate([np.zeros_like(g), np.ones_like(g), yg, yg]))


def grid_points(G):
    """Interior G x G grid + boundary ring in the layout of the reference's generator (data.py:192-217)."""
    g = np.linspace(0.0, 1.0, G + 2)
    xi, yi = np.meshgrid(g[1:-1], g[1:-1])
    xb, yb = ring(G)
    return xi.flatten(), yi.flatten(), xb, yb


def perlin(x, y, seed=0):
    """One Perlin-family source on the host (mirror of the device generator)."""
    from fast_poisson_solver_b200.sources import perlin_params, perlin_sources_numpy

    return perlin_sources_numpy(x, y, perlin_params(1, seed=seed))[:, 0]


# ---------------------------------------------------------------------------------------------
# clocks sampler (profiling recipe)
# ---------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", os.environ.get("FPS_BENCH_CLOCK_MS", "100")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        busy = [v for v in sm if v > 500] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's algorithm-faithful port on host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(n_sample, seed=0, laplace="autograd"):
    """One bounded step of the reference algorithm on the CPU: precompute (model.h + 2 800 autograd
    sweeps of utils.py:54-68 + Gram + LHS) and one solve, fp32 like the reference's default."""
    import torch
    from oracle import fps_oracle as o

    w = o.init_weights(0)
    x, y = sobol_points(n_sample, seed)
    xb, yb = ring(max(2, int(round(n_sample ** 0.5))))
    f = perlin(x, y, seed)
    bc = np.full_like(xb, 3.0)
    t0 = time.perf_counter()
    st = o.precompute(x, y, xb, yb, w, dtype=np.float32, laplace=laplace)
    o.solve(st, f, bc)
    return time.perf_counter() - t0, torch.get_num_threads()


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 128
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    for _ in range(args.warmup):
        cpu_reference_step(n_sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, threads = cpu_reference_step(n_sample)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = n_sample / dt
    sample = (f"each step = {n_sample} of the 2^20 scattered points + their ring, precompute+solve, oracle port of the "
              f"reference algorithm (model.h + 2800 torch.autograd sweeps, fp32) on {threads} host threads; the reference "
              f"is pure Python and cannot travel to the GPU box, its per-point cost is size-independent beyond ~1e3 points "
              f"(SURVEY.md 6: 844 s for 1e4 points on 8 cores)")
    print(json.dumps({
        "impl": "reference", "metric": "precompute_points_per_s", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.gpus, args.points),
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
def measure_peak(torch, dtype, n, tf32=False, reps=10):
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    best = 1e30
    try:
        for _ in range(3):
            a @ b
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        burst = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        # sustained: back to back for ~1.5 s under the power cap (MEASURED_PEAKS.json protocol)
        reps_s = max(10, int(1.5 / (best * 1e-3)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps_s):
            a @ b
        e1.record()
        e1.synchronize()
        sustained = 2.0 * n ** 3 * reps_s / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return burst, sustained


def measure_int8_peak(torch, n=8192, reps=10):
    """Dense int8 tensor-core rate (cuBLASLt through torch._int_mm), TOP/s: best of `reps` and back to back for ~1.5 s."""
    try:
        a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda").t().contiguous().t()
        for _ in range(3):
            torch._int_mm(a, b)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps_s = max(10, int(1.5 / (best * 1e-3)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps_s):
            torch._int_mm(a, b)
        e1.record()
        e1.synchronize()
        return {"burst_tops": 2.0 * n ** 3 / (best * 1e-3) / 1e12,
                "sustained_tops": 2.0 * n ** 3 * reps_s / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": f"torch._int_mm {n}^3 (cuBLASLt int8 -> int32), best of {reps} / back to back for 1.5 s, this run"}
    except Exception as exc:  # noqa: BLE001
        return {"burst_tops": None, "sustained_tops": None, "how": f"torch._int_mm unavailable: {exc}"}


def read_profile(solver):
    from fast_poisson_solver_b200 import _lib

    out = {}
    for k, name in enumerate(_lib.PROF_KINDS):
        ms, n = C.c_double(), C.c_int64()
        _lib.check(solver._lib.fps_profile_read(solver._ctx, k, C.byref(ms), C.byref(n)), "fps_profile_read")
        out[name] = (ms.value, n.value)
    return out


def ncu_traffic(capture, points_per_launch=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full
    capture of this round (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py), scaled to the points one launch
    of THIS run covers; None when no capture is committed."""
    for tag in ("r02b", "r02"):            # r02b: re-capture of the layer kernel after the segment / pitch changes
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json")))
            for k, v in t["bytes_per_launch"].items():
                if k.startswith(capture + ":"):
                    return v * (points_per_launch / t["points_per_layer_launch"] if points_per_launch else 1.0)
        except Exception:
            pass
    return None


def multi_gpu_parity(torch, dist, world, rank, local):
    """A point-sharded precompute + batched and single solve over all ranks against the SAME problem solved on rank 0
    alone (multi-GPU parity must be driver-visible).  20 736 points per rank: the exact int8 paths."""
    from fast_poisson_solver_b200 import Solver
    from fast_poisson_solver_b200.parallel import shard_bounds

    G = int(round((20736 * world) ** 0.5))
    xp, yp, xb, yb = grid_points(G)
    B = 40
    Fm = np.stack([(50.0 + j) * np.sin((2 + j % 4) * np.pi * xp) * np.sin((3 + (j // 4) % 3) * np.pi * yp) for j in range(B)], 1)
    Fm[:, 1] = perlin(xp, yp, seed=3)
    bcv = np.linspace(-2, 3, B)
    lo, hi = shard_bounds(xp.size, rank, world)
    blo, bhi = shard_bounds(xb.size, rank, world)
    s = Solver(device=f"cuda:{local}", use_weights=False, distributed=True)
    s.precompute(xp[lo:hi], yp[lo:hi], xb[blo:bhi], yb[blo:bhi])
    u_pde = s.solve(Fm[lo:hi], np.tile(bcv[None], (bhi - blo, 1)))[1]
    u1 = s.solve(Fm[lo:hi, 1], np.full(bhi - blo, bcv[1]))[1]
    sizes = [shard_bounds(xp.size, r, world)[1] - shard_bounds(xp.size, r, world)[0] for r in range(world)]
    mine = torch.zeros((max(sizes), B + 1), dtype=torch.float32, device=f"cuda:{local}")   # equal sizes for all_gather
    mine[: hi - lo] = torch.cat([u_pde, u1], dim=1)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    out = None
    if rank == 0:
        full = torch.cat([g[:n] for g, n in zip(gathered, sizes)], dim=0).double()
        ref = Solver(device=f"cuda:{local}", use_weights=False)
        # the calibration (fp16 operand scales, fixed-point grid of DH) belongs to the problem: the sharded run combined
        # it over the ranks, the single-GPU run takes the same state -> bit-identical features on both sides
        ref.set_calibration(s.calibration_state())
        ref.precompute(xp, yp, xb, yb)
        r_pde = ref.solve(Fm, bcv)[1].double()
        r1 = ref.solve(Fm[:, 1], np.full(xb.size, bcv[1]))[1].double()
        eb = max(float((full[:, j] - r_pde[:, j]).norm() / r_pde[:, j].norm()) for j in range(B))
        e1 = float((full[:, B] - r1[:, 0]).norm() / r1[:, 0].norm())
        out = {"points": int(xp.size), "rhs": B, "ranks": world, "worst_u_rel_l2_batched": eb, "u_rel_l2_single": e1,
               "tolerance": 5e-6, "pass": bool(eb < 5e-6 and e1 < 5e-6),
               "note": "same kernels, same calibration state (combined over the ranks with one small all-reduce), hence "
                       "bit-identical features; what differs is the summation order of the fp64 partial Grams / projections, "
                       "which the 1e5 : 1 cancellation of a rough source (column 1, Perlin) amplifies to ~1e-6; the smooth "
                       "columns agree to < 1e-6 (tests/test_gpu_multi.py: 8 ranks, 6.8e-7 smooth, 1.7e-6 rough)"}
        del ref
    del s
    torch.cuda.empty_cache()
    dist.barrier()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a collective that waits longer than this is a dead-lock, not a slow rank: fail in minutes, not in NCCL's default 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=300))
    from fast_poisson_solver_b200 import Solver, _lib

    os.makedirs("/tmp/fps_bench_cwd", exist_ok=True)
    os.chdir("/tmp/fps_bench_cwd")
    n_pts = args.points
    dev = torch.device("cuda", local)
    x, y = sobol_points(n_pts, seed=rank)           # every rank owns its own points (weak scaling)
    xb, yb = ring(GRID)
    nb = xb.size // world
    xb, yb = xb[rank * nb:(rank + 1) * nb if rank < world - 1 else None], yb[rank * nb:(rank + 1) * nb if rank < world - 1 else None]
    f = perlin(x, y, seed=rank).astype(np.float32)
    bc = np.full(xb.size, 3.0, dtype=np.float32)

    solver = Solver(device=f"cuda:{local}", use_weights=False, seed=0, distributed=world > 1, engine=args.engine)
    lib, ctx = solver._lib, solver._ctx

    d = lambda a, dt: torch.as_tensor(a, dtype=dt, device=dev)
    xd, yd, xbd, ybd = d(x, torch.float64), d(y, torch.float64), d(xb, torch.float64), d(yb, torch.float64)
    fd, bcd = d(f, torch.float32), d(bc, torch.float32)

    def step_resident():
        solver.precompute(xd, yd, xbd, ybd)
        return solver.solve(fd, bcd)

    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hx, hy, hxb, hyb, hf, hbc = pin(x), pin(y), pin(xb), pin(yb), pin(f), pin(bc)
    hu = torch.empty((n_pts + xb.size, 1), dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (hx, hy, hxb, hyb, hf, hbc))
    d2h = hu.numel() * 4

    def step_e2e():
        a = [t.to(dev, non_blocking=True) for t in (hx, hy, hxb, hyb, hf, hbc)]
        solver.precompute(a[0], a[1], a[2], a[3])
        u = solver.solve(a[4], a[5])[0]
        hu.copy_(u, non_blocking=True)
        torch.cuda.synchronize()
        return hu

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warm = max(3, args.warmup)
    # the clock sampler (an nvidia-smi child process) starts BEFORE the warm-up: its NVML initialisation would
    # otherwise fall into the first timed steps
    clocks = Clocks(local)
    if rank == 0 and os.environ.get("FPS_BENCH_NO_CLOCKS") != "1":
        clocks.start()
        time.sleep(1.0)
    for _ in range(warm):
        step_resident()
    # ---- headline: K steps, instrumentation OFF --------------------------------------------------
    if rank == 0:
        clocks.rows.clear()
    launches0 = lib.fps_launch_count()
    ms_step = timed(step_resident, args.steps)
    launches = (lib.fps_launch_count() - launches0) // args.steps
    clk = clocks.stop() if rank == 0 else None
    # ---- the same steps again with per-kernel CUDA events (roofline of the dominant kernel, kernel breakdown) -----
    _lib.check(lib.fps_profile_enable(ctx, 1), "profile")
    _lib.check(lib.fps_profile_reset(ctx), "profile")
    ms_step_prof = timed(step_resident, args.steps)
    prof = read_profile(solver)
    _lib.check(lib.fps_profile_enable(ctx, 0), "profile")

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    cal = (C.c_double * 5)()
    lib.fps_get_calibration(ctx, cal)
    planes_cached = solver._gplanes is not None

    # ---- batched solve (configs[2]): 512x512 grid, B right-hand sides PER RANK ---------------------------
    batched = None
    if args.batch > 0:
        from bench_configs import mixed_sources
        from fast_poisson_solver_b200.sources import grid_points as dev_grid_points

        solver_b = Solver(device=f"cuda:{local}", use_weights=False, engine=args.engine)
        gx, gy, gxb, gyb = dev_grid_points(solver_b, 512)
        solver_b.precompute(gx, gy, gxb, gyb)
        B = args.batch
        Fm = mixed_sources(solver_b, B, seed=10 + rank)                     # Perlin + sine families, generated on the device
        bcv = (torch.rand(B, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)) * 20 - 10)
        solver_b.solve(Fm, bcv)
        ms_b = timed(lambda: solver_b.solve(Fm, bcv), 3)
        _lib.check(lib.fps_profile_enable(solver_b._ctx, 1), "profile")
        _lib.check(lib.fps_profile_reset(solver_b._ctx), "profile")
        timed(lambda: solver_b.solve(Fm, bcv), 3)
        pb = read_profile(solver_b)
        _lib.check(lib.fps_profile_enable(solver_b._ctx, 0), "profile")
        u_cols = solver_b.u_pde_pred[:, [0, B - 1]].double()
        f_cols, bc_cols = Fm[:, [0, B - 1]].double(), bcv[[0, B - 1]].double().cpu().numpy()
        npts_b = gx.numel()
        batched = {"value": world * B / (ms_b * 1e-3), "unit": "solves/s", "ms_per_call": ms_b,
                   "config": {"workload": "configs[2]: 4096 Perlin/sine RHS on a 512x512 grid", "points": int(npts_b), "rhs": B,
                              "sharding": "B columns per rank, features replicated, no collective (strong form: `strong` block)"},
                   "planes_cached": {"gram": solver_b._gplanes is not None, "recon": isinstance(solver_b._rplanes_h, torch.Tensor)},
                   "kernels_ms": {k: round(v[0] / 3, 3) for k, v in pb.items() if v[1]},
                   "issued_int8_tops": {
                       # digit-pair products actually issued: projection 16 pairs (4 x 4 digits), reconstruction 20 (5 x 4)
                       "project": 16 * 2.0 * 768 * ((B + 63) // 64 * 64) * npts_b / (pb["project"][0] / 3 * 1e-3) / 1e12 if pb["project"][0] else None,
                       "reconstruct": 2 * 20 * 2.0 * 704 * ((B + 255) // 256 * 256) * npts_b / (pb["reconstruct"][0] / 3 * 1e-3) / 1e12 if pb["reconstruct"][0] else None}}
        # accuracy oracle of north_star: the reference's finite-difference solve on the same 512x512 grid (GPU DST, pinned
        # to the reference's numeric_solve by tests/golden/numeric_solve_seed0.npz), metrics as analyze.py:107-119
        from fast_poisson_solver_b200.numeric import fd_solve_grid

        u_fd = fd_solve_grid(solver_b, 512, 512, 1.0 / 513, 1.0 / 513, f_cols.contiguous(), bc=bc_cols)
        acc = []
        for j, fam in enumerate(("perlin", "sine")):
            dlt = u_cols[:, j] - u_fd[:, j]
            mae = float(dlt.abs().mean())
            acc.append({"family": fam, "mse": float((dlt * dlt).mean()), "rmse": float((dlt * dlt).mean().sqrt()), "mae": mae,
                        "rmae": mae / max(float(u_fd[:, j].abs().mean()), 1e-6), "rel_l2": float(dlt.norm() / u_fd[:, j].norm())})
        batched["accuracy_vs_fd_512x512"] = acc
        del Fm, solver_b
        torch.cuda.empty_cache()

    parity = multi_gpu_parity(torch, dist, world, rank, local) if world > 1 else None

    # ---- peaks (rank 0 measures; the other ranks idle) ------------------------------------------------------
    peaks, tf32_burst, tf32_peak, f64_burst, f64_peak, i8 = {}, None, None, None, None, None
    if rank == 0:
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf32_burst, tf32_peak = measure_peak(torch, torch.float32, 8192, tf32=True)
        f64_burst, f64_peak = measure_peak(torch, torch.float64, 4096)
        i8 = measure_int8_peak(torch)
    barrier()

    # ---- strong-scaling target configurations --------------------------------------------------------------
    strong = None
    if args.strong:
        from bench_configs import Env, strong_block

        del solver
        torch.cuda.empty_cache()
        try:
            strong = strong_block(Env(), quick=args.quick)
        except Exception as e:      # the headline line must survive a failure of the extra configurations
            import traceback

            traceback.print_exc()
            strong = {"scaling": "strong", "n_gpus": world, "error": f"{type(e).__name__}: {e}"[:500]}
            torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel ----------------------------------------------------------
    # the FP16x2-split engine issues kind::f16 MMAs: the tensor roofline is the dense 16-bit peak of MEASURED_PEAKS.json
    # (bf16 and fp16 share the rate); sustained figure, the kernel runs inside a long step
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    tc_burst = float(peaks.get("bf16_tflops", 1590.0))
    tc_src = ("dense 16-bit tensor peak from MEASURED_PEAKS.json (bf16_tflops_sustained; fp16 MMAs run at the same rate)"
              if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)")
    hid_ms, hid_n = prof["hidden"]
    hid_flop = 4 * args.steps * n_pts * FLOP_PER_POINT_HIDDEN      # 4 hidden launches per wave write planes
    achieved = hid_flop / (hid_ms * 1e-3) / 1e12 if hid_ms else None
    kernels_ms = {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]}
    feat_ms = sum(prof[k][0] for k in ("fourier", "layer0", "hidden", "last")) / args.steps
    gram_ms = prof["gram"][0] / args.steps
    gram_flop = 700 * 701 * (n_pts + xb.size)
    gram_issued = (16 * 23 * 256 * 64 * 2 * ((n_pts + 127) // 128 * 128)) / (gram_ms * 1e-3) / 1e12 if gram_ms else None
    i8_sus = i8.get("sustained_tops") if i8 else None
    solve_ms = sum(prof[k][0] for k in ("project", "trisolve", "reconstruct")) / args.steps
    out = {
        "metric": "precompute_points_per_s", "value": world * n_pts / (ms_step * 1e-3), "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16x2-split(3 MMA) features + exact int8 (4-digit) Gram + f64 factor/solve",
        "data": "synthetic",
        "config": bench_config(world, n_pts),
        "e2e": {"value": world * n_pts / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clk,
        "engine": {"features": "tcgen05" if solver_engine_is_tc(args) else "simt", "gram_planes_cached": planes_cached,
                   "accumulator_gain_check": {"H_with": cal[0], "DH_with": cal[1], "H_without": cal[2], "DH_without": cal[3],
                                              "gain_in_use": cal[4]}},
        "ms_per_step_instrumented": ms_step_prof,
        "kernels_ms_per_step": kernels_ms,
        "roofline": {"kernel": "layer_tc_pair_kernel<4,false,true,128> (hidden 700x700 layer, 4 streams, FP16x2-split tcgen05 cta_group::2)",
                     "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": achieved / tc_peak if achieved else None,
                     "issued_frac": 3 * achieved * (768 / 700) * (704 / 700) / tc_peak if achieved else None,
                     "frac_ceiling": 1.0 / (3 * (768 / 700) * (704 / 700)),
                     "peak_burst": tc_burst, "tf32_peak_sustained_this_run": tf32_peak,
                     "peak_source": tc_src + "; achieved = ALGORITHMIC flops (2*4*700*700 per point) / CUDA-event time of the "
                                    "instrumented pass; fp32-grade products need 3 MMAs each and M is padded 700->768, so frac "
                                    "cannot exceed frac_ceiling; issued_frac is the tensor-pipe share actually occupied; the "
                                    "kernel also sits at the L2->SM operand-feed cap (DESIGN.md section 4)",
                     "launches": int(hid_n), "avg_launch_ms": hid_ms / hid_n if hid_n else None,
                     "points_per_launch": 4 * args.steps * n_pts / hid_n if hid_n else None,
                     "traffic": ncu_traffic("hidden", 4 * args.steps * n_pts / hid_n if hid_n else None),
                     "traffic_unit": "bytes/launch: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                                     "(profiles/r02b_ncu_traffic.json); null = no capture committed for this build"},
        "int8_peak": i8,
        "other_rooflines": {
            "features_all_layers": {"achieved": n_pts * FLOP_PER_POINT_FEATURES / (feat_ms * 1e-3) / 1e12,
                                    "peak": tc_peak, "unit": "TFLOP/s (algorithmic, 3 MMAs issued per product)"},
            # exact Gram on the int8 tensor cores (csrc/gram_i8.cu): 16 digit-pair products per fp64-grade product
            "gram_exact_int8": {"achieved": gram_flop / (gram_ms * 1e-3) / 1e12 if gram_ms else None,
                                "unit": "TFLOP/s (algorithmic SYRK flops, fp64-exact result)",
                                "fp64_pipe_peak": f64_peak, "fp64_pipe_peak_burst": f64_burst,
                                "issued_int8_tops": gram_issued,
                                "issued_frac_of_measured_int8_peak": gram_issued / i8_sus if gram_issued and i8_sus else None,
                                "note": "issued int8 work = 16 digit pairs x 23 tiles of 256 x 64 x 2 ops per point against the "
                                        "int8 peak measured in this run; the digit split is fused into the last layer's epilogue"},
            "single_solve_hbm": {"ms": solve_ms, "algorithmic_bytes": 8400 * n_pts,
                                 "achieved_gbs": 8400 * n_pts / (solve_ms * 1e-3) / 1e9 if solve_ms else None,
                                 "peak_gbs": float(peaks.get("hbm_gbs", 6545.6)),
                                 "frac": 8400 * n_pts / (solve_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6545.6)) if solve_ms else None,
                                 "note": "projection + triangular sweeps + reconstruction of one right-hand side; 8400 B/point = "
                                         "DH twice + H once (SURVEY.md 8a8)"},
        },
        "batched_solve": batched,
        "multi_gpu_parity": parity,
        "strong": strong,
    }
    if batched and i8_sus:
        for k in ("project", "reconstruct"):
            v = batched["issued_int8_tops"].get(k)
            batched["issued_int8_tops"][k + "_frac_of_measured_peak"] = v / i8_sus if v else None
    # ---- CPU baseline (bounded sample, rank 0, N = 1 only) ----------------------------------------
    if world == 1 and not args.no_cpu:
        n_sample = 512
        dt, threads = cpu_reference_step(n_sample)
        dt_fwd, _ = cpu_reference_step(10000, laplace="forward")
        out["cpu_baseline"] = {
            "value": n_sample / dt, "unit": "points/s", "cores": threads, "kind": "port",
            "sample": f"{n_sample} of the 2^20 points: oracle port of the reference algorithm (model.h + 2800 autograd "
                      f"sweeps, fp32, precompute+solve) took {dt:.1f} s on {threads} threads of {os.cpu_count()} cores",
            "forward_mode_port_points_per_s": 10000 / dt_fwd,
        }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def solver_engine_is_tc(args):
    return args.engine == "tcgen05"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--engine", default="tcgen05")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", dest="strong", action="store_false")
    ap.add_argument("--quick", action="store_true", help="reduced sizes in the strong block (development)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
