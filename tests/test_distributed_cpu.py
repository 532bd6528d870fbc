"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU path (row sharding + the packed
all-reduce) reproduces the single-process normal equations.  The per-shard feature blocks come from
the CPU oracle here; on the GPU box the same PackedNormal / allreduce code runs over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fps_oracle as o


def _worker(rank, world, port, G, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fast_poisson_solver_b200.parallel import PackedNormal, allreduce_sum, shard_bounds

        w = o.init_weights(0)
        xp, yp, xb, yb = o.grid_problem(G)
        lo, hi = shard_bounds(xp.size, rank, world)
        blo, bhi = shard_bounds(xb.size, rank, world)
        H, DH = o.features_forward(xp[lo:hi], yp[lo:hi], w, np.float64)
        Hb, _ = o.features_forward(xb[blo:bhi], yb[blo:bhi], w, np.float64, laplace=False)
        pk = PackedNormal("cpu")
        pk.G[:700, :700] = torch.from_numpy(DH.T @ DH)
        pk.Gbc[:700, :700] = torch.from_numpy(Hb.T @ Hb)
        pk.s[:700] = torch.from_numpy(Hb.sum(0))
        pk.set_counts(hi - lo, bhi - blo)
        pk.allreduce()
        N, Nbc = pk.counts()
        f = 50.0 * np.sin(2 * np.pi * xp) * np.sin(3 * np.pi * yp)
        R = torch.from_numpy(DH.T @ f[lo:hi].reshape(-1, 1))
        allreduce_sum(R)
        sum_bc = torch.tensor([3.0 * (bhi - blo)], dtype=torch.float64)
        allreduce_sum(sum_bc)
        # replicated solve, row-sharded reconstruction
        lam = o.DEFAULT_LAMBDA
        G64, Gbc, s = pk.G[:700, :700].numpy(), pk.Gbc[:700, :700].numpy(), pk.s[:700].numpy()
        LHS = lam / N * G64 + (Gbc - np.outer(s, s) / Nbc) / Nbc
        wv = np.linalg.solve(LHS, lam / N * R.numpy())
        bias = -(s @ wv - sum_bc.numpy()) / Nbc
        u_loc = H @ wv + bias
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), u=u_loc, lo=lo, hi=hi, N=N, Nbc=Nbc, LHS=LHS)
    finally:
        dist.destroy_process_group()


def test_row_sharded_normal_equations_match_single_process(tmp_path):
    G, world = 32, 2  # N + N_bc = 1156 > 700 unknowns: u is well defined
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, G, str(tmp_path)), nprocs=world, join=True)
    w = o.init_weights(0)
    xp, yp, xb, yb = o.grid_problem(G)
    st = o.precompute(xp, yp, xb, yb, w)
    f = 50.0 * np.sin(2 * np.pi * xp) * np.sin(3 * np.pi * yp)
    u_ref = o.solve(st, f, np.full_like(xb, 3.0))[1]
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert int(parts[0]["N"]) == xp.size and int(parts[0]["Nbc"]) == xb.size
    assert np.array_equal(parts[0]["LHS"], parts[1]["LHS"])          # replicated state is bit-identical
    assert np.allclose(parts[0]["LHS"], st.LHSs[0], rtol=1e-11, atol=1e-16)
    u = np.concatenate([p["u"] for p in parts])
    assert [int(p["lo"]) for p in parts] == [0, xp.size // 2]
    assert o.rel_l2(u, u_ref) < 1e-6


def test_shard_bounds_cover_everything():
    from fast_poisson_solver_b200.parallel import shard_bounds

    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
