"""GPU, >= 2 devices: row-sharded precompute + solve over NCCL against the single-GPU result."""
import os

import numpy as np
import pytest
import torch

from oracle import fps_oracle as o

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, G, B, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    import datetime

    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    try:
        from fast_poisson_solver_b200 import Solver
        from fast_poisson_solver_b200.parallel import shard_bounds

        os.chdir(out_dir)
        xp, yp, xb, yb = o.grid_problem(G)
        lo, hi = shard_bounds(xp.size, rank, world)
        blo, bhi = shard_bounds(xb.size, rank, world)
        Fm = np.stack([(50.0 + j) * np.sin((2 + j % 4) * np.pi * xp) * np.sin((3 + (j // 4) % 3) * np.pi * yp) for j in range(B)], 1)
        bcv = np.linspace(-2, 3, B)
        s = Solver(device=f"cuda:{rank}", use_weights=False, distributed=True)
        s.precompute(xp[lo:hi], yp[lo:hi], xb[blo:bhi], yb[blo:bhi])
        assert s.NO == xp.size and s.NdO == xb.size
        u, u_pde, u_bc, f_pred, _ = s.solve(Fm[lo:hi], np.tile(bcv[None], (bhi - blo, 1)))
        u1 = s.solve(Fm[lo:hi, 0], np.full(bhi - blo, bcv[0]))[1]
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), u_pde=u_pde.cpu().numpy(), u_bc=u_bc.cpu().numpy(),
                 u1=u1.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2])
@pytest.mark.parametrize("G,B", [(48, 5), (192, 40)])   # (192, 40): 18 432 points per rank -> the exact int8 Gram /
def test_row_sharded_matches_single_gpu(tmp_path, world, G, B):   # projection / reconstruction paths on every rank
    import torch.multiprocessing as mp
    from fast_poisson_solver_b200 import Solver

    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, G, B, str(tmp_path)), nprocs=world, join=True)
    xp, yp, xb, yb = o.grid_problem(G)
    Fm = np.stack([(50.0 + j) * np.sin((2 + j % 4) * np.pi * xp) * np.sin((3 + (j // 4) % 3) * np.pi * yp) for j in range(B)], 1)
    bcv = np.linspace(-2, 3, B)
    s = Solver(use_weights=False)
    s.precompute(xp, yp, xb, yb)
    u, u_pde, u_bc, _, _ = s.solve(Fm, bcv)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    up = np.concatenate([p["u_pde"] for p in parts])
    ub = np.concatenate([p["u_bc"] for p in parts])
    # same kernels, same features; the summation order of the Gram / projection differs (fp64) and, on the int8 path,
    # every rank snaps its rows of DH to the grid of its own column maxima (2^-31 of the maximum)
    assert o.rel_l2(up, u_pde.cpu().numpy()) < 1e-6
    assert o.rel_l2(ub, u_bc.cpu().numpy()) < 1e-6
    assert o.rel_l2(np.concatenate([p["u1"] for p in parts]), u_pde[:, [0]].cpu().numpy()) < 1e-6
