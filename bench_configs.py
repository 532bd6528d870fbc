#!/usr/bin/env python
"""The BASELINE target configurations beyond bench.py's headline step (STRONG scaling: the total problem is fixed, ranks
own row shards of the points or column shards of the right-hand sides).  bench.py imports `strong_block` and puts its
result on the driver's JSON line; the CLI runs one configuration alone.

    python bench_configs.py --config c3|c4|c5 [...]
    torchrun --nproc-per-node N ... bench_configs.py --config c4|c5        (points sharded over N GPUs)

c3  BASELINE configs[2]: 4096 Perlin/sine right-hand sides on a 512x512 grid, COLUMNS sharded over the ranks (features
    replicated, no collective on the data path).
c4  BASELINE configs[3]: irregular domain (L-shape [0,1]^2 minus (0.5,1]^2, or the disc r = 0.5), scrambled Sobol points
    rejected to the domain, boundary points uniformly along the perimeter; precompute with point sharding + one NCCL
    all-reduce of the packed Gram buffer; accuracy = residual metrics of analyze.py:107-119 (f_pred vs f, u_bc vs bc) -
    the FD oracle needs a tensor grid (numeric_solver.py:95-104).
c5  BASELINE configs[4]: 4096^2 grid (16.7 M points, generated on the device) + 4096 batched solves, Perlin and sine
    sources generated on the device in column blocks, outputs reduced to checksums, accuracy of the first columns
    against the reference's finite-difference solve on a 2048^2 grid (GPU DST, csrc/fd_dst.cu, pinned to the reference's
    numeric_solve by tests/golden/numeric_solve_seed0.npz) through Solver.evaluate.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def domain_points(domain, n, seed=0):
    from scipy.stats import qmc

    eng = qmc.Sobol(d=2, scramble=True, seed=seed)
    xs, ys, have = [], [], 0
    while have < n:
        s = eng.random(1 << 20)
        x, y = s[:, 0], s[:, 1]
        keep = ~((x > 0.5) & (y > 0.5)) if domain == "lshape" else ((x - 0.5) ** 2 + (y - 0.5) ** 2 < 0.25)
        xs.append(x[keep]); ys.append(y[keep]); have += int(keep.sum())
    x, y = np.concatenate(xs)[:n], np.concatenate(ys)[:n]
    nb = int(4 * np.sqrt(n))
    t = (np.arange(nb) + 0.5) / nb
    if domain == "disc":
        xb, yb = 0.5 + 0.5 * np.cos(2 * np.pi * t), 0.5 + 0.5 * np.sin(2 * np.pi * t)
        xb, yb = np.clip(xb, 0, 1), np.clip(yb, 0, 1)
    else:  # L-shape perimeter, length 4: (0,0)->(1,0)->(1,.5)->(.5,.5)->(.5,1)->(0,1)->(0,0)
        corners = np.array([[0, 0], [1, 0], [1, .5], [.5, .5], [.5, 1], [0, 1], [0, 0]], dtype=np.float64)
        seg = np.linalg.norm(np.diff(corners, axis=0), axis=1)
        cum = np.concatenate([[0], np.cumsum(seg)])
        d = t * cum[-1]
        k = np.clip(np.searchsorted(cum, d, side="right") - 1, 0, len(seg) - 1)
        a = (d - cum[k]) / seg[k]
        xb = corners[k, 0] + a * (corners[k + 1, 0] - corners[k, 0])
        yb = corners[k, 1] + a * (corners[k + 1, 1] - corners[k, 1])
    return x, y, xb, yb


class Env:
    """Rank / device / timing helpers shared by the configurations."""

    def __init__(self):
        import torch

        self.torch = torch
        self.world, self.rank, self.local = (int(os.environ.get(k, d)) for k, d in
                                             (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps=1):
        """Device time of `reps` calls (CUDA events on the current stream), max over ranks, ms per call."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / reps)


def mixed_sources(solver, B, seed, out=None):
    """B right-hand sides on the solver's points, generated on the device: first half Perlin family, second half sine
    family (data_utils/source_functions.py:173-227, :24-73)."""
    import torch
    from fast_poisson_solver_b200.sources import perlin_params, perlin_sources, sine_params, sine_sources

    n = solver._xy64[0].numel()
    if out is None:
        out = torch.empty((n, B), dtype=solver._fdtype, device=solver._tdev)
    nper = (B + 1) // 2
    perlin_sources(solver, perlin_params(nper, seed=seed), out=out[:, :nper])
    if B > nper:
        sine_sources(solver, sine_params(B - nper, seed=seed + 1), out=out[:, nper:B])
    return out


def run_c3(env, solver=None, grid=512, rhs=4096, reps=3):
    """configs[2]: `rhs` right-hand sides on a grid x grid domain, columns sharded across the ranks."""
    import torch
    from fast_poisson_solver_b200 import Solver
    from fast_poisson_solver_b200.parallel import shard_bounds
    from fast_poisson_solver_b200.sources import grid_points

    s = solver or Solver(device=f"cuda:{env.local}", use_weights=False)
    x, y, xb, yb = grid_points(s, grid)
    s.precompute(x, y, xb, yb)
    lo, hi = shard_bounds(rhs, env.rank, env.world)
    Fm = mixed_sources(s, hi - lo, seed=100 + env.rank)
    bcv = torch.as_tensor(np.random.default_rng(env.rank).uniform(-10, 10, hi - lo), device=env.dev)
    s.solve(Fm, bcv)
    ms = env.timed(lambda: s.solve(Fm, bcv), reps)
    return {"workload": f"configs[2]: {rhs} Perlin/sine RHS on a {grid}x{grid} grid, columns sharded x{env.world}",
            "points": grid * grid, "rhs": rhs, "rhs_per_rank": hi - lo, "ms_per_call": ms,
            "solves_per_s": rhs / (ms * 1e-3), "collectives_on_data_path": 0}, s


def run_c4(env, domain="lshape", points=4194304, reps=2):
    """configs[3]: irregular domain, point-sharded precompute + one packed all-reduce, one solve."""
    import torch
    import torch.distributed as dist
    from fast_poisson_solver_b200 import Solver
    from fast_poisson_solver_b200.parallel import shard_bounds
    from fast_poisson_solver_b200.sources import perlin_params, perlin_sources

    x, y, xb, yb = domain_points(domain, points)
    N, Nb = x.size, xb.size
    lo, hi = shard_bounds(N, env.rank, env.world)
    blo, bhi = shard_bounds(Nb, env.rank, env.world)
    d = lambda v: torch.as_tensor(v, dtype=torch.float64, device=env.dev)
    xd, yd, xbd, ybd = d(x[lo:hi]), d(y[lo:hi]), d(xb[blo:bhi]), d(yb[blo:bhi])
    s = Solver(device=f"cuda:{env.local}", use_weights=False, distributed=env.world > 1)
    for _ in range(2):                      # the host-side point generation let the GPU clocks fall: warm up twice
        s.precompute(xd, yd, xbd, ybd)
    ms_pre = env.timed(lambda: s.precompute(xd, yd, xbd, ybd), reps)
    f = perlin_sources(s, perlin_params(1, seed=5))
    bc = torch.full((bhi - blo,), 2.5, device=env.dev)
    for _ in range(2):
        s.solve(f, bc)
    ms_solve = env.timed(lambda: s.solve(f, bc), max(reps, 6))     # ms-scale calls: a single host hiccup must not dominate
    u, u_pde, u_bc, f_pred, _ = s.solve(f, bc)
    # residual metrics (analyze.py:107-119), reduced over ranks
    acc = torch.stack([((f_pred - f) ** 2).sum().double(), (f ** 2).sum().double(), (f_pred - f).abs().sum().double(),
                       ((u_bc - 2.5) ** 2).sum().double(), (u_bc - 2.5).abs().sum().double()])
    if env.world > 1:
        dist.all_reduce(acc)
    acc = acc.cpu().numpy()
    out = {"workload": f"configs[3]: {domain} domain, {N} scattered points sharded x{env.world}, Perlin source",
           "points": N, "boundary_points": Nb, "precompute_ms": ms_pre, "precompute_points_per_s": N / (ms_pre * 1e-3),
           "solve_ms": ms_solve, "jitter": s.jitter, "f_residual_rel_l2": float(np.sqrt(acc[0] / acc[1])),
           "f_mae": float(acc[2] / N), "bc_rmse": float(np.sqrt(acc[3] / Nb)), "bc_mae": float(acc[4] / Nb),
           "collectives": "1 all-reduce (7.9 MB fp64) per precompute, 1 per solve"}
    del s
    torch.cuda.empty_cache()
    return out


def run_c5(env, grid=4096, rhs=4096, block=None, fd_grid=2048, ncheck=2):
    """configs[4]: grid^2 points sharded over the ranks + `rhs` batched solves in column blocks; FD accuracy check."""
    import torch
    import torch.distributed as dist
    from fast_poisson_solver_b200 import Solver
    from fast_poisson_solver_b200.numeric import fd_solve_grid
    from fast_poisson_solver_b200.sources import grid_points

    s = Solver(device=f"cuda:{env.local}", use_weights=False, distributed=env.world > 1)
    x, y, xb, yb = grid_points(s, grid, env.rank, env.world)
    N, Nb, n_loc = grid * grid, 4 * grid + 4, x.numel()
    s.precompute(x, y, xb, yb)
    ms_pre = env.timed(lambda: s.precompute(x, y, xb, yb), 1)
    if block is None:
        # per column of a block: F, U and f_pred rows (fp32) + the digit planes of F: ~13 bytes per point; use 60 % of what is free
        free, _ = torch.cuda.mem_get_info(env.local)
        free += torch.cuda.memory_reserved(env.local) - torch.cuda.memory_allocated(env.local)
        block = int(max(64, min(512, int(free * 0.6) // (n_loc * 13) // 64 * 64)))
    bcv = np.random.default_rng(1).uniform(-10, 10, rhs)
    Fbuf = torch.empty((n_loc, block), dtype=torch.float32, device=env.dev)
    chk = torch.zeros(2, dtype=torch.float64, device=env.dev)

    def all_solves(record=False, first_only=False):
        for c0 in range(0, rhs, block):
            c1 = min(rhs, c0 + block)
            Fm = mixed_sources(s, c1 - c0, seed=1000 + c0, out=Fbuf[:, : c1 - c0])
            u = s.solve(Fm, torch.as_tensor(bcv[c0:c1], device=env.dev))[0]
            if record:
                # fp64 accumulation without an fp64 copy of u
                for j0 in range(0, u.shape[1], 16):
                    part = u[:, j0:j0 + 16].double()
                    chk[0] += part.sum()
                    chk[1] += (part * part).sum()
                    del part
            if first_only:
                break

    all_solves(first_only=True)          # warm-up: workspaces, reconstruction planes
    ms_solves = env.timed(lambda: all_solves(record=False), 1)
    # source generation is part of the loop above; time it alone to report the solver's own share
    def gen_only():
        for c0 in range(0, rhs, block):
            mixed_sources(s, min(rhs, c0 + block) - c0, seed=1000 + c0, out=Fbuf[:, : min(rhs, c0 + block) - c0])
    ms_gen = env.timed(gen_only, 1)
    all_solves(record=True, first_only=True)
    if env.world > 1:
        dist.all_reduce(chk)
    out = {"workload": f"configs[4]: {grid}x{grid} grid ({N} points) sharded x{env.world} + {rhs} Perlin/sine solves",
           "points": N, "boundary_points": Nb, "rhs": rhs, "column_block": block,
           "planes_cached": {"gram": s._gplanes is not None, "recon": isinstance(s._rplanes_h, torch.Tensor)},
           "precompute_ms": ms_pre, "precompute_points_per_s": N / (ms_pre * 1e-3),
           "solves_ms_total": ms_solves, "source_generation_ms": ms_gen, "solves_per_s": rhs / ((ms_solves - ms_gen) * 1e-3),
           "solves_per_s_incl_source_generation": rhs / (ms_solves * 1e-3), "jitter": s.jitter,
           "checksum_first_block": [float(chk[0].item()), float(chk[1].item())]}
    # accuracy of the first columns vs the reference's FD solve on an fd_grid^2 grid: u evaluated there (rank 0) against
    # the GPU DST solve of the same sources
    Fm = mixed_sources(s, ncheck, seed=1000, out=Fbuf[:, :ncheck])
    s.solve(Fm, torch.as_tensor(bcv[:ncheck], device=env.dev))
    if env.rank == 0:
        from fast_poisson_solver_b200.sources import (perlin_params, perlin_sources, sine_params, sine_sources)

        gx, gy, _, _ = grid_points(s, fd_grid)
        u_ml = s.evaluate(gx, gy).double()
        Ffd = torch.empty((fd_grid * fd_grid, ncheck), dtype=torch.float64, device=env.dev)
        nper = (ncheck + 1) // 2
        perlin_sources(s, perlin_params(nper, seed=1000), x=gx, y=gy, out=Ffd[:, :nper])
        if ncheck > nper:
            sine_sources(s, sine_params(ncheck - nper, seed=1001), x=gx, y=gy, out=Ffd[:, nper:])
        h = 1.0 / (fd_grid + 1)
        u_fd = fd_solve_grid(s, fd_grid, fd_grid, h, h, Ffd, bc=bcv[:ncheck])
        cols = []
        for j in range(ncheck):
            dlt = u_ml[:, j] - u_fd[:, j]
            mae = float(dlt.abs().mean())
            cols.append({"family": "perlin" if j < nper else "sine", "mse": float((dlt * dlt).mean()),
                         "rmse": float((dlt * dlt).mean().sqrt()), "mae": mae,
                         "rmae": mae / max(float(u_fd[:, j].abs().mean()), 1e-6),
                         "rel_l2": float(dlt.norm() / u_fd[:, j].norm())})
        out["accuracy_vs_fd"] = {"fd_grid": fd_grid, "solver": "fps_fd_solve (GPU DST-I, fp64)", "columns": cols}
    env.barrier()
    del s, Fbuf
    torch.cuda.empty_cache()
    return out


def strong_block(env, quick=False):
    """What bench.py adds to the driver's line: the three target configurations at this GPU count."""
    out = {"scaling": "strong", "n_gpus": env.world}
    c3, s3 = run_c3(env, grid=512, rhs=4096)
    out["configs[2]_rhs_sharded"] = c3
    del s3
    env.torch.cuda.empty_cache()
    out["configs[3]_lshape_4M"] = run_c4(env, "lshape", 4194304 if not quick else 1 << 19)
    out["configs[4]_16M_4096_solves"] = run_c5(env, grid=4096 if not quick else 1024, rhs=4096 if not quick else 512,
                                               fd_grid=2048 if not quick else 512)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c3", "c4", "c5", "all"])
    ap.add_argument("--domain", default="lshape", choices=["lshape", "disc"])
    ap.add_argument("--points", type=int, default=4194304)
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--rhs", type=int, default=4096)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--fd-grid", type=int, default=2048)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    env = Env()
    torch.cuda.set_device(env.local)
    if env.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        dist.init_process_group("nccl", device_id=env.dev, timeout=datetime.timedelta(seconds=300))
    os.makedirs("/tmp/fps_bench_cwd", exist_ok=True)
    os.chdir("/tmp/fps_bench_cwd")
    if a.config == "c3":
        out, _ = run_c3(env, grid=512 if a.grid == 4096 else a.grid, rhs=a.rhs)
    elif a.config == "c4":
        out = run_c4(env, a.domain, a.points)
    elif a.config == "c5":
        out = run_c5(env, a.grid, a.rhs, a.block or None, a.fd_grid)
    else:
        out = strong_block(env, quick=a.quick)
    out["n_gpus"] = env.world
    if env.world > 1:
        dist.barrier()
    if env.rank == 0:
        print(json.dumps(out), flush=True)
    if env.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
