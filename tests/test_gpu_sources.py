"""GPU: parity on BASELINE's own source families against goldens minted from the unmodified reference, the device
source / geometry generators against their host mirrors, and the GPU finite-difference solver against the reference's
own numeric_solve."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fps_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def make_solver():
    from fast_poisson_solver_b200 import Solver

    made = []

    def make(**kw):
        kw.setdefault("use_weights", False)
        s = Solver(**kw)
        made.append(s)
        return s

    yield make
    made.clear()


@pytest.mark.parametrize("name", ["sources_G32_seed0.npz", "sources_G100_seed0.npz"])
def test_baseline_source_families_vs_reference_golden(golden, make_solver, name):
    """VERDICT r1 parity gap (a).  Six draws of the reference's own sine family (k ~ U(-10,10)), four Perlin-family fields
    and the benign control, solved by the UNMODIFIED reference in fp64 and fp32 (oracle/make_golden.py sources).
      precision=float32:  err(ours, ref64) <= max(1e-5, err(ref32, ref64))  per source - the reference's own fp32 run is
                          the yardstick where feature rounding is amplified by |w| ~ 1e4 (north_star: fp32 path);
      precision=float64:  err(ours, ref64) <= 1e-5 on every source.
    Single solves and one batched solve of all columns."""
    if not os.path.exists(os.path.join(GOLDEN, name)):
        pytest.skip(f"{name} not minted")
    g = golden(name)
    xp, yp, xb, yb = o.grid_problem(int(g["G"]))
    names = [str(n) for n in g["names"]]
    F, bcv, U64, U32 = g["F"], g["bc"], g["U_f64"], g["U_f32"]
    floor = [o.rel_l2(U32[:, j], U64[:, j]) for j in range(len(names))]
    for prec in (torch.float32, torch.float64):
        s = make_solver(precision=prec)
        s.precompute(xp, yp, xb, yb)
        ub = s.solve(F, bcv)[0].double().cpu().numpy()
        rows = []
        for j, nm in enumerate(names):
            u1 = s.solve(F[:, j], np.full(xb.size, bcv[j]))[0].double().cpu().numpy()
            e1, eb = o.rel_l2(u1[:, 0], U64[:, j]), o.rel_l2(ub[:, j], U64[:, j])
            rows.append((nm, e1, eb, floor[j]))
            bound = 1e-5 if prec == torch.float64 else max(1e-5, floor[j])
            assert e1 <= bound and eb <= bound, (name, str(prec), nm, e1, eb, floor[j])
        print(f"\n[{name} {prec}] u rel-L2 vs reference fp64 (single | batched | reference fp32 vs fp64):")
        for nm, e1, eb, fl in rows:
            print(f"   {nm:8s} {e1:.2e} | {eb:.2e} | {fl:.2e}")


def test_device_perlin_and_grid_generators(make_solver):
    """fps_perlin_sources / fps_grid_points against their host mirrors (sources.py, data.py:192-217 layout)."""
    from fast_poisson_solver_b200.sources import grid_points, perlin_params, perlin_sources, perlin_sources_numpy

    s = make_solver()
    G = 37
    xp, yp, xb, yb = o.grid_problem(G)
    x, y, xbd, ybd = grid_points(s, G)
    assert np.array_equal(x.cpu().numpy(), xp) and np.array_equal(y.cpu().numpy(), yp)
    assert np.array_equal(xbd.cpu().numpy(), xb) and np.array_equal(ybd.cpu().numpy(), yb)
    # shards concatenate to the whole
    parts = [grid_points(s, G, r, 3) for r in range(3)]
    for k, ref in enumerate((xp, yp, xb, yb)):
        assert np.array_equal(torch.cat([p[k] for p in parts]).cpu().numpy(), ref)
    s.precompute(x, y, xbd, ybd)
    par = perlin_params(45, seed=9)
    Fd = perlin_sources(s, par)
    Fh = perlin_sources_numpy(xp, yp, par)
    assert Fd.shape == (G * G, 45) and Fd.dtype == torch.float32
    assert np.allclose(Fd.cpu().numpy(), Fh, rtol=1e-6, atol=1e-4 * np.abs(Fh).max())
    rng = np.random.default_rng(0)
    xs, ys = rng.uniform(0, 1, 1000), rng.uniform(0, 1, 1000)
    s64 = make_solver(precision=torch.float64)
    s64.precompute(xs, ys, xb, yb)
    F64 = perlin_sources(s64, par)
    assert F64.dtype == torch.float64
    Fh = perlin_sources_numpy(xs, ys, par)
    assert np.abs(F64.cpu().numpy() - Fh).max() < 1e-9 * max(1.0, np.abs(Fh).max())


def test_gpu_fd_solver_vs_reference_numeric_solve(golden, make_solver):
    """fps_fd_solve (DST-I as fp64 GEMMs) against the reference's own numeric_solve (golden from oracle/make_golden.py
    numeric) and against the CPU oracle; drop-in numeric_solve() call incl. shuffled points and non-constant boundary data."""
    from fast_poisson_solver_b200 import numeric_solve
    from fast_poisson_solver_b200.numeric import fd_solve_grid
    from oracle import fd_oracle as fd

    g = golden("numeric_solve_seed0.npz")
    s = make_solver()
    for G in (48, 200):
        xp, yp, xb, yb = o.grid_problem(G)
        for tag in ("sine", "smooth"):
            f, bcv, u_ref = g[f"f_G{G}_{tag}"], float(g[f"bc_G{G}_{tag}"]), g[f"u_G{G}_{tag}"]
            u, dt = numeric_solve(f, xp, yp, np.full(xb.size, bcv), xb, yb, precision=torch.float64, solver=s)
            assert isinstance(u, np.ndarray) and u.shape == u_ref.shape and u.dtype == np.float64 and dt > 0
            assert o.rel_l2(u, u_ref) < 1e-10, (G, tag)
    # the README's 400 x 400 case (5 s in the reference): analytic source, golden stored at stride 7
    G = 400
    xp, yp, xb, yb = o.grid_problem(G)
    f = 50.0 * np.sin(2 * np.pi * xp) * np.sin(3 * np.pi * yp)
    u, dt = numeric_solve(f, xp, yp, np.full(xb.size, 3.0), xb, yb, precision=torch.float64, solver=s)
    assert o.rel_l2(u[::7], g["u_G400_smooth"]) < 1e-9
    u32, _ = numeric_solve(f, xp, yp, np.full(xb.size, 3.0), xb, yb, solver=s)
    assert u32.dtype == np.float32 and o.rel_l2(u32, u) < 1e-6
    # shuffled point order + non-constant boundary values vs the sparse-LU restatement of the same system
    G = 24
    xp, yp, xb, yb = o.grid_problem(G)
    rng = np.random.default_rng(1)
    f = o.sine_source(xp, yp, seed=7)
    ub_vals = 1.0 + xb - 2.0 * yb * yb
    pi, pb = rng.permutation(xp.size), rng.permutation(xb.size)
    u, _ = numeric_solve(f[pi], xp[pi], yp[pi], ub_vals[pb], xb[pb], yb[pb], precision=torch.float64, solver=s)
    import scipy.sparse as sp
    from scipy.sparse.linalg import spsolve

    n = G + 2
    h = 1.0 / (G + 1)
    d2 = sp.diags([1.0, -2.0, 1.0], [-1, 0, 1], shape=(n, n))
    Lm = ((sp.kron(sp.eye(n), d2) + sp.kron(d2, sp.eye(n))) / h ** 2).tolil()
    idx = np.arange(n * n).reshape(n, n)
    rhs = np.zeros(n * n)
    gx = np.rint(xp / h).astype(int); gy = np.rint(yp / h).astype(int)
    rhs[idx[gy, gx]] = f
    bx = np.rint(xb / h).astype(int); by = np.rint(yb / h).astype(int)
    for k in idx[by, bx]:
        Lm.rows[k] = [int(k)]; Lm.data[k] = [1.0]
    rhs[idx[by, bx]] = ub_vals
    full = spsolve(Lm.tocsr(), rhs)
    ref = np.concatenate([full[idx[gy, gx]][pi], full[idx[by, bx]][pb]])
    assert o.rel_l2(u, ref) < 1e-10
    # batched (groups of 2), fp32 right-hand sides, vs the CPU DST oracle
    G = 130
    xp, yp, _, _ = o.grid_problem(G)
    F = np.stack([o.sine_source(xp, yp, seed=j) for j in range(5)], 1).astype(np.float32)
    bcs = np.array([0.0, 1.0, -2.0, 3.0, 0.5])
    U = fd_solve_grid(s, G, G, 1.0 / (G + 1), 1.0 / (G + 1), torch.tensor(F, device="cuda"), bc=bcs, batch=2)
    assert o.rel_l2(U.cpu().numpy(), fd.fd_solve_dst(G, F.astype(np.float64), bcs)) < 1e-12
    with pytest.raises(ValueError):
        numeric_solve(f[:-1], xp[:-1], yp[:-1], np.zeros(xb.size), xb, yb, solver=s)
