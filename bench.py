#!/usr/bin/env python
"""Benchmark of the fast-poisson-solver hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...            (one rank per GPU, NCCL)

One step = one pass of the hot path over one batch of synthetic input:
    Solver.precompute(x_pde, y_pde, x_bc, y_bc)  +  Solver.solve(f, u_bc)
on BASELINE.json configs[1]: 1024x1024 scattered (scrambled Sobol) collocation points in the unit
square, 4100 boundary-ring points, Perlin source, constant Dirichlet value.  With N > 1 GPUs every rank
owns its own 2^20 points (weak scaling); the Gram / boundary blocks are all-reduced once per
precompute and the projected right-hand side once per solve.

Prints ONE JSON line (rank 0).  `value` = points/s with inputs resident in HBM; `e2e` = the same step
fed from pinned host buffers with the result copied back; `roofline` = the dominant kernel (hidden
700x700 layer, tcgen05 3xTF32) timed live with CUDA events; `cpu_baseline` = the oracle's
algorithm-faithful port of the reference (2 800 autograd sweeps) on this box's host cores;
`batched_solve` = BASELINE configs[2] (4096 right-hand sides on a 512x512 grid).
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


# ---------------------------------------------------------------------------------------------
# synthetic inputs
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


def perlin(x, y, seed=0, layers=3):
    """Vectorised gradient noise, statistically like source_functions.py:173-227 (1-5 octave layers,
    octaves ~ U(0.1,12), amplitude ~ U(-100,100)); values need not match the absent perlin_noise
    package - sources are inputs to the path."""
    rng = np.random.default_rng(seed)
    out = np.zeros_like(x)
    for _ in range(layers):
        octv = rng.uniform(0.1, 12.0)
        amp = rng.uniform(-100, 100)
        n = int(np.ceil(octv)) + 2
        ang = rng.uniform(0, 2 * np.pi, (n + 1, n + 1))
        gx, gy = np.cos(ang), np.sin(ang)
        px, py = x * octv, y * octv
        ix, iy = np.floor(px).astype(np.int64), np.floor(py).astype(np.int64)
        fx, fy = px - ix, py - iy
        sx, sy = fx * fx * fx * (fx * (fx * 6 - 15) + 10), fy * fy * fy * (fy * (fy * 6 - 15) + 10)

        def dot(ax, ay, dx, dy):
            return gx[ax, ay] * dx + gy[ax, ay] * dy

        n00 = dot(ix, iy, fx, fy)
        n10 = dot(ix + 1, iy, fx - 1, fy)
        n01 = dot(ix, iy + 1, fx, fy - 1)
        n11 = dot(ix + 1, iy + 1, fx - 1, fy - 1)
        out += amp * ((n00 * (1 - sx) + n10 * sx) * (1 - sy) + (n01 * (1 - sx) + n11 * sx) * sy)
    return out


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
    sample = (f"{n_sample} of the 2^20 scattered points + their ring, precompute+solve, oracle port of the reference "
              f"algorithm (model.h + 2800 torch.autograd sweeps, fp32) on host cores")
    print(json.dumps({
        "impl": "reference", "metric": "precompute_points_per_s", "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 1024x1024 scattered points, Perlin source, precompute + single solve",
                   "sample": sample},
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
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    burst = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    # sustained: back to back for ~1.5 s under the power cap (MEASURED_PEAKS.json protocol)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    try:
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


def read_profile(solver):
    from fast_poisson_solver_b200 import _lib

    out = {}
    for k, name in enumerate(_lib.PROF_KINDS):
        ms, n = C.c_double(), C.c_int64()
        _lib.check(solver._lib.fps_profile_read(solver._ctx, k, C.byref(ms), C.byref(n)), "fps_profile_read")
        out[name] = (ms.value, n.value)
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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

    for _ in range(max(3, args.warmup)):
        step_resident()
    _lib.check(lib.fps_profile_enable(ctx, 1), "profile")
    _lib.check(lib.fps_profile_reset(ctx), "profile")
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    launches0 = lib.fps_launch_count()
    ms_step = timed(step_resident, args.steps)
    launches = (lib.fps_launch_count() - launches0) // args.steps
    prof = read_profile(solver)
    _lib.check(lib.fps_profile_enable(ctx, 0), "profile")
    clk = clocks.stop() if rank == 0 else None

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- batched solve (configs[2]): 512x512 grid, B right-hand sides ---------------------------
    batched = None
    if args.batch > 0:
        gx, gy, gxb, gyb = grid_points(512)
        solver.precompute(d(gx, torch.float64), d(gy, torch.float64), d(gxb, torch.float64), d(gyb, torch.float64))
        B = args.batch
        gen = torch.Generator(device=dev).manual_seed(rank)
        base = d(np.stack([perlin(gx, gy, seed=s) for s in range(8)], 1), torch.float32)       # (n, 8)
        mix = torch.randn(8, B, device=dev, generator=gen)
        Fm = (base @ mix).contiguous()                                                          # (n, B) synthetic mixtures
        bcv = (torch.rand(B, device=dev, generator=gen) * 20 - 10)
        solver.solve(Fm, bcv)
        _lib.check(lib.fps_profile_enable(ctx, 1), "profile")
        _lib.check(lib.fps_profile_reset(ctx), "profile")
        ms_b = timed(lambda: solver.solve(Fm, bcv), 3)
        u_cols = solver.u_pde_pred[:, :4].double().cpu().numpy()
        f_cols, bc_cols = Fm[:, :4].double().cpu().numpy(), bcv[:4].double().cpu().numpy()
        pb = read_profile(solver)
        _lib.check(lib.fps_profile_enable(ctx, 0), "profile")
        batched = {"value": world * B / (ms_b * 1e-3), "unit": "solves/s", "ms_per_call": ms_b,
                   "config": {"workload": "configs[2]: 4096 RHS on a 512x512 grid", "points": int(gx.size), "rhs": B,
                              "sharding": "columns across ranks, features replicated, no collective"},
                   "kernels_ms": {k: round(v[0] / 3, 3) for k, v in pb.items() if v[1]}}
        del Fm

    if world > 1:
        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf32_burst, tf32_peak = measure_peak(torch, torch.float32, 8192, tf32=True)
    f16_fmt = os.environ.get("FPS_TC_F16", "1") != "0"
    if f16_fmt:
        # the FP16x2-split engine issues kind::f16 MMAs: the tensor roofline is the dense 16-bit peak of
        # MEASURED_PEAKS.json (bf16 and fp16 share the rate); sustained figure, the kernel runs inside a long step
        tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        tc_burst = float(peaks.get("bf16_tflops", 1590.0))
        tc_src = ("dense 16-bit tensor peak from MEASURED_PEAKS.json (bf16_tflops_sustained; fp16 MMAs run at the same rate)"
                  if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)")
    else:
        tc_peak, tc_burst = tf32_peak, tf32_burst
        tc_src = "dense TF32 measured in this run (torch.matmul 8192^3, allow_tf32, sustained 1.5 s)"
    f64_burst, f64_peak = measure_peak(torch, torch.float64, 4096)
    hid_ms, hid_n = prof["hidden"]
    hid_flop = 4 * args.steps * n_pts * FLOP_PER_POINT_HIDDEN      # 4 hidden launches per chunk write planes
    achieved = hid_flop / (hid_ms * 1e-3) / 1e12 if hid_ms else None
    kernels_ms = {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]}
    feat_ms = sum(prof[k][0] for k in ("fourier", "layer0", "hidden", "last")) / args.steps
    gram_ms = prof["gram"][0] / args.steps
    gram_flop = 700 * 701 * (n_pts + xb.size)
    out = {
        "metric": "precompute_points_per_s", "value": world * n_pts / (ms_step * 1e-3), "unit": "points/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16x2-split(3 MMA) features + exact int8 (4-digit) Gram + f64 factor/solve",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 1024x1024 scattered (Sobol) points, Perlin source, precompute + single solve",
                   "points_per_gpu": n_pts, "boundary_points": int(xb.size) * world, "weights": "default init, seed 0",
                   "engine": "tcgen05" if solver.engine == 1 else "simt", "lambda": 2 ** -12,
                   "parallelism": f"points sharded x{world}, 1 all-reduce/precompute + 1/solve" if world > 1 else "single GPU",
                   "l2": "inputs (H, DH: 2.9 GB each) larger than L2; no flush needed"},
        "e2e": {"value": world * n_pts / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clk,
        "kernels_ms_per_step": kernels_ms,
        "roofline": {"kernel": "layer_tc_pair_kernel<4,false,true> (hidden 700x700 layer, 4 streams, FP16x2-split tcgen05 cta_group::2)",
                     "bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                     "frac": achieved / tc_peak if achieved else None,
                     "issued_frac": 3 * achieved * (768 / 700) * (704 / 700) / tc_peak if achieved else None,
                     "frac_ceiling": 1.0 / (3 * (768 / 700) * (704 / 700)),
                     "peak_burst": tc_burst, "tf32_peak_sustained_this_run": tf32_peak,
                     "peak_source": tc_src + "; achieved = ALGORITHMIC flops (2*4*700*700 per point) / CUDA-event time; "
                                    "fp32-grade products need 3 MMAs each and M is padded 700->768, so frac cannot "
                                    "exceed frac_ceiling; issued_frac is the tensor-pipe share actually occupied",
                     "launches": int(hid_n), "avg_launch_ms": hid_ms / hid_n if hid_n else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture
                     # (profiles/r01d_layer_tc_ncu_full_summary.csv: 390.8 MB for an 18 944-point launch = 20.6 KB per
                     # point; algorithmic: 11.3 KB in + 11.3 KB out per point + 2 MB of weights), scaled to the points
                     # of one launch of this run
                     "traffic": (20.6e3 if f16_fmt else 44.7e3) * (4 * args.steps * n_pts / hid_n) if hid_n else None,
                     "traffic_unit": "bytes/launch (ncu, per-point figure x points per launch)"},
        "other_rooflines": {
            "features_all_layers": {"achieved": n_pts * FLOP_PER_POINT_FEATURES / (feat_ms * 1e-3) / 1e12,
                                    "peak": tc_peak, "unit": "TFLOP/s (algorithmic, 3 MMAs issued per product)"},
            # exact Gram on the int8 tensor cores (csrc/gram_i8.cu): 16 digit-pair products per fp64-grade product
            "gram_exact_int8": {"achieved": gram_flop / (gram_ms * 1e-3) / 1e12 if gram_ms else None,
                                "unit": "TFLOP/s (algorithmic SYRK flops, fp64-exact result)",
                                "fp64_pipe_peak": f64_peak, "fp64_pipe_peak_burst": f64_burst,
                                "peak_source": "the fp64 pipe this path replaces: torch.matmul fp64 4096^3 sustained / "
                                               "best of 10, this run; issued int8 work = 16 digit pairs x 23 tiles of 256 x 64 x 2 ops per point, "
                                               "against a nominal 4.5 POP/s dense int8 peak; the time includes the digit split",
                                "issued_int8_tops": (16 * 23 * 256 * 64 * 2 * ((n_pts + 127) // 128 * 128)) / (gram_ms * 1e-3) / 1e12
                                if gram_ms else None},
        },
        "batched_solve": batched,
    }
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
        if batched is not None:
            # accuracy oracle of north_star: the reference's finite-difference solve (numeric_solver.py:39-160,
            # restated exactly by a DST in oracle/fd_oracle.py) on the 512x512 grid of the batched config,
            # metrics as analyze.py:107-119
            from oracle import fd_oracle, fps_oracle

            u_fd = fd_oracle.fd_solve_dst(512, f_cols, bc_cols)
            out["cpu_baseline"]["accuracy_vs_fd_512x512"] = [
                dict(fps_oracle.metrics(u_cols[:, j], u_fd[:, j]), rel_l2=fps_oracle.rel_l2(u_cols[:, j], u_fd[:, j]))
                for j in range(u_cols.shape[1])]
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
