"""Row-sharded precompute + solve against the single-GPU result: over NCCL with one rank per device (>= 2 devices), and
with 8 ranks sharing ONE device over gloo (the same host logic and kernels; runs on a single-GPU box as well)."""
import os

import numpy as np
import pytest
import torch

from oracle import fps_oracle as o

pytestmark = pytest.mark.gpu


def _sources(xp, yp, B):
    """Smooth product sines, plus - column 1 - a rough field in the style of BASELINE's sine family (twelve terms,
    wave numbers up to 10: |w| ~ 1e5 and u comes out of a 1e5 : 1 cancellation, the hard case for any reordering)."""
    Fm = np.stack([(50.0 + j) * np.sin((2 + j % 4) * np.pi * xp) * np.sin((3 + (j // 4) % 3) * np.pi * yp) for j in range(B)], 1)
    rng = np.random.default_rng(7)
    rough = np.zeros_like(xp)
    for _ in range(12):
        k1, k2, a = rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-100, 100)
        rough += a * np.sin(k1 * np.pi * (xp - rng.uniform())) * np.cos(k2 * np.pi * (yp - rng.uniform()))
    if B > 1:
        Fm[:, 1] = rough
    return Fm, np.linspace(-2, 3, B)


def _worker(rank, world, port, G, B, out_dir, shared_gpu=False):
    import datetime

    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if shared_gpu else rank
    torch.cuda.set_device(dev)
    if shared_gpu:
        # NCCL cannot place two ranks on one device: gloo carries the same all-reduces (FPS_COMM=torch keeps the
        # library's own NCCL communicator out of the way)
        os.environ["FPS_COMM"] = "torch"
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                                timeout=datetime.timedelta(seconds=120))
    try:
        from fast_poisson_solver_b200 import Solver
        from fast_poisson_solver_b200.parallel import shard_bounds

        os.chdir(out_dir)
        xp, yp, xb, yb = o.grid_problem(G)
        lo, hi = shard_bounds(xp.size, rank, world)
        blo, bhi = shard_bounds(xb.size, rank, world)
        Fm, bcv = _sources(xp, yp, B)
        s = Solver(device=f"cuda:{dev}", use_weights=False, distributed=True)
        s.precompute(xp[lo:hi], yp[lo:hi], xb[blo:bhi], yb[blo:bhi])
        assert s.NO == xp.size and s.NdO == xb.size
        u, u_pde, u_bc, f_pred, _ = s.solve(Fm[lo:hi], np.tile(bcv[None], (bhi - blo, 1)))
        u1 = s.solve(Fm[lo:hi, 0], np.full(bhi - blo, bcv[0]))[1]
        cal = s.calibration_state()      # combined over the ranks: identical everywhere
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), u_pde=u_pde.cpu().numpy(), u_bc=u_bc.cpu().numpy(),
                 u1=u1.cpu().numpy(), scales=np.asarray(cal["scales"]), gain_on=cal["gain_on"], dh_cmax=cal["dh_cmax"])
    finally:
        dist.destroy_process_group()


def _check_against_single_gpu(tmp_path, world, G, B, tol, tol_rough=1e-5):
    from fast_poisson_solver_b200 import Solver

    xp, yp, xb, yb = o.grid_problem(G)
    Fm, bcv = _sources(xp, yp, B)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for p in parts[1:]:          # every rank ended up with the same calibration
        assert np.array_equal(p["scales"], parts[0]["scales"]) and np.array_equal(p["dh_cmax"], parts[0]["dh_cmax"])
        assert int(p["gain_on"]) == int(parts[0]["gain_on"])
    s = Solver(use_weights=False)
    # the calibration belongs to the problem: with the sharded run's state the single-GPU features are bit-identical,
    # what remains is the summation order of the fp64 partial Grams / projections
    s.set_calibration({"scales": parts[0]["scales"], "gain_on": int(parts[0]["gain_on"]), "dh_cmax": parts[0]["dh_cmax"]})
    s.precompute(xp, yp, xb, yb)
    u, u_pde, u_bc, _, _ = s.solve(Fm, bcv)
    up = np.concatenate([p["u_pde"] for p in parts])
    ub = np.concatenate([p["u_bc"] for p in parts])
    ref = u_pde.cpu().numpy()
    col = [o.rel_l2(up[:, [j]], ref[:, [j]]) for j in range(B)]
    smooth = max(c for j, c in enumerate(col) if j != 1)
    errs = (smooth, o.rel_l2(ub, u_bc.cpu().numpy()),
            o.rel_l2(np.concatenate([p["u1"] for p in parts]), ref[:, [0]]))
    print(f"world {world} G {G} B {B}: smooth columns {errs[0]:.2e} u_bc {errs[1]:.2e} single {errs[2]:.2e} "
          f"rough column {col[1] if B > 1 else 0.0:.2e}")
    assert max(errs) < tol, errs
    if B > 1:
        assert col[1] < tol_rough, col[1]


@pytest.mark.parametrize("world,G,B", [(8, 407, 40)])   # 20 706 points per rank: the exact int8 paths on every rank,
def test_eight_ranks_on_one_gpu_match_single_solver(tmp_path, world, G, B):   # the sizes of bench.py's parity block
    import torch.multiprocessing as mp

    port = 29800 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, G, B, str(tmp_path), True), nprocs=world, join=True)
    _check_against_single_gpu(tmp_path, world, G, B, 2e-6)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2])
@pytest.mark.parametrize("G,B", [(48, 5), (192, 40)])   # (192, 40): 18 432 points per rank -> the exact int8 Gram /
def test_row_sharded_matches_single_gpu(tmp_path, world, G, B):   # projection / reconstruction paths on every rank
    import torch.multiprocessing as mp

    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, G, B, str(tmp_path)), nprocs=world, join=True)
    _check_against_single_gpu(tmp_path, world, G, B, 2e-6)
