#!/usr/bin/env python
"""Large BASELINE configurations that do not fit bench.py's one-line contract.

    python bench_configs.py --config c4 [--domain lshape|disc] [--points 4194304]
    python bench_configs.py --config c5 [--grid 4096] [--rhs 4096] [--block 512]
    torchrun --nproc-per-node N ... bench_configs.py --config c4|c5        (points sharded over N GPUs)

c4  BASELINE configs[3]: irregular domain (L-shape [0,1]^2 minus (0.5,1]^2, or the disc r = 0.5), scrambled
    Sobol points rejected to the domain, boundary points uniformly along the perimeter; precompute with
    point sharding + one NCCL all-reduce of the packed Gram buffer; accuracy = residual metrics of
    analyze.py (f_pred vs f, u_bc vs bc) - the FD oracle needs a tensor grid (numeric_solver.py:95-104).
c5  BASELINE configs[4]: 4096^2 grid (16.7 M points) + 4096 batched solves, sources generated on the
    device in column blocks, outputs reduced to checksums, accuracy of the first columns against the
    reference's FD solve (exact DST restatement) on a 2048^2 grid through Solver.evaluate.
Prints one JSON object (rank 0).  STRONG scaling: the total problem is fixed, ranks own row shards.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c4", "c5"])
    ap.add_argument("--domain", default="lshape", choices=["lshape", "disc"])
    ap.add_argument("--points", type=int, default=4194304)
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--rhs", type=int, default=4096)
    ap.add_argument("--block", type=int, default=512)
    ap.add_argument("--fd-grid", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from fast_poisson_solver_b200 import Solver
    from fast_poisson_solver_b200.parallel import shard_bounds
    from fast_poisson_solver_b200.sources import sine_params, sine_sources, sine_sources_numpy
    from bench import grid_points

    os.makedirs("/tmp/fps_bench_cwd", exist_ok=True)
    os.chdir("/tmp/fps_bench_cwd")

    def sync_max(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return sync_max(e0.elapsed_time(e1) / reps)

    if a.config == "c4":
        x, y, xb, yb = domain_points(a.domain, a.points)
        N, Nb = x.size, xb.size
    else:
        x, y, xb, yb = grid_points(a.grid)
        N, Nb = x.size, xb.size
    lo, hi = shard_bounds(N, rank, world)
    blo, bhi = shard_bounds(Nb, rank, world)
    d = lambda v: torch.as_tensor(v, dtype=torch.float64, device=dev)
    xd, yd, xbd, ybd = d(x[lo:hi]), d(y[lo:hi]), d(xb[blo:bhi]), d(yb[blo:bhi])
    s = Solver(device=f"cuda:{local}", use_weights=False, distributed=world > 1)
    s.precompute(xd, yd, xbd, ybd)
    ms_pre = timed(lambda: s.precompute(xd, yd, xbd, ybd), a.reps)
    out = {"config": a.config, "n_gpus": world, "points": N, "boundary_points": Nb, "scaling": "strong",
           "precompute_ms": ms_pre, "precompute_points_per_s": N / (ms_pre * 1e-3), "jitter": s.jitter}

    if a.config == "c4":
        out["domain"] = a.domain
        par = sine_params(1, seed=5)
        f = sine_sources(s, par)
        bc = torch.full((bhi - blo,), 2.5, device=dev)
        s.solve(f, bc)
        ms_solve = timed(lambda: s.solve(f, bc), a.reps)
        u, u_pde, u_bc, f_pred, _ = s.solve(f, bc)
        # residual metrics (analyze.py:107-119), reduced over ranks
        acc = torch.stack([((f_pred - f) ** 2).sum().double(), (f ** 2).sum().double(), (f_pred - f).abs().sum().double(),
                           ((u_bc - 2.5) ** 2).sum().double(), (u_bc - 2.5).abs().sum().double()])
        if world > 1:
            dist.all_reduce(acc)
        acc = acc.cpu().numpy()
        out.update({"solve_ms": ms_solve, "f_residual_rel_l2": float(np.sqrt(acc[0] / acc[1])),
                    "f_mae": float(acc[2] / N), "bc_rmse": float(np.sqrt(acc[3] / Nb)), "bc_mae": float(acc[4] / Nb)})
    else:
        B, blk = a.rhs, a.block
        par = sine_params(B, seed=0)
        bcv = np.random.default_rng(1).uniform(-10, 10, B)
        Fbuf = torch.empty((hi - lo, blk), dtype=torch.float32, device=dev)
        chk = torch.zeros(2, dtype=torch.float64, device=dev)

        def all_solves(record=False):
            for c0 in range(0, B, blk):
                c1 = min(B, c0 + blk)
                Fm = sine_sources(s, par[c0:c1], out=Fbuf[:, : c1 - c0])
                u = s.solve(Fm, torch.as_tensor(bcv[c0:c1], device=dev))[0]
                if record:
                    # fp64 accumulation without an fp64 copy of u (34 GB per 256 columns at 16.7 M points)
                    for j0 in range(0, u.shape[1], 16):
                        part = u[:, j0:j0 + 16].double()
                        chk[0] += part.sum()
                        chk[1] += (part * part).sum()
                        del part

        all_solves()
        ms_solves = timed(all_solves, 1)
        all_solves(record=True)
        if world > 1:
            dist.all_reduce(chk)
        out.update({"rhs": B, "column_block": blk, "solves_ms_total": ms_solves, "solves_per_s": B / (ms_solves * 1e-3),
                    "checksum_sum_u": float(chk[0].item()), "checksum_sum_u2": float(chk[1].item())})
        # accuracy of the first columns vs the reference's FD solve on a fd_grid^2 grid (rank 0 evaluates u there)
        ncheck = 2
        Fm = sine_sources(s, par[:ncheck], out=Fbuf[:, :ncheck])
        s.solve(Fm, torch.as_tensor(bcv[:ncheck], device=dev))
        if rank == 0:
            # checker leg: the oracle (test infrastructure) is imported only here, to judge accuracy
            from oracle import fd_oracle, fps_oracle as o

            G = a.fd_grid
            gx, gy, _, _ = grid_points(G)
            u_ml = s.evaluate(d(gx), d(gy)).double().cpu().numpy()
            u_fd = fd_oracle.fd_solve_dst(G, sine_sources_numpy(gx, gy, par[:ncheck]), bcv[:ncheck])
            out["accuracy_vs_fd"] = {"fd_grid": G, "columns": [dict(o.metrics(u_ml[:, j], u_fd[:, j]),
                                                                    rel_l2=o.rel_l2(u_ml[:, j], u_fd[:, j])) for j in range(ncheck)]}
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
