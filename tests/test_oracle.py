"""CPU: the oracle restatement against golden vectors minted from the unmodified reference
(oracle/make_golden.py).  These tests are what pins the oracle."""
import hashlib

import numpy as np
import pytest

from oracle import fps_oracle as o


def _digest(w):
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()


def test_default_init_matches_reference_pinn(golden, weights0):
    g = golden("features_seed0.npz")
    assert _digest(weights0) == str(g["weights_sha256"])


def test_features_forward_vs_reference_autograd_fp64(golden, weights0):
    g = golden("features_seed0.npz")
    H, DH = o.features_forward(g["x"], g["y"], weights0, np.float64)
    assert o.rel_l2(H, g["H_f64"]) < 1e-13
    assert o.rel_l2(DH, g["DH_f64"]) < 1e-13


def test_features_forward_fp32_noise_floor(golden, weights0):
    g = golden("features_seed0.npz")
    H, DH = o.features_forward(g["x"], g["y"], weights0, np.float32)
    assert o.rel_l2(H, g["H_f64"]) < 3e-6
    assert o.rel_l2(DH, g["DH_f64"]) < 3e-6


def test_autograd_port_matches_reference(golden, weights0):
    g = golden("features_seed0.npz")
    H, DH = o.features_autograd(g["x"][:4], g["y"][:4], weights0, np.float64)
    assert o.rel_l2(H, g["H_f64"][:4]) < 1e-13
    assert o.rel_l2(DH, g["DH_f64"][:4]) < 1e-13


@pytest.mark.parametrize("name", ["solve_G32_seed0.npz", "solve_G100_seed0.npz"])
def test_precompute_solve_vs_reference_fp64(golden, weights0, name):
    import os
    from conftest import GOLDEN

    if not os.path.exists(os.path.join(GOLDEN, name)):
        pytest.skip(f"{name} not minted")
    g = golden(name)
    G = int(g["G"])
    xp, yp, xb, yb = o.grid_problem(G)
    st = o.precompute(xp, yp, xb, yb, weights0, dtype=np.float64)
    u, _, _, f_pred, _, bias = o.solve(st, g["f"], g["bc"])
    assert o.rel_l2(st.DHtDH[::7, ::7], g["DHtDH_s7_f64"]) < 1e-12
    assert o.rel_l2(st.LHSs[0][::7, ::7], g["LHS_s7_f64"]) < 1e-10
    assert o.rel_l2(st.Ht_bc_ones, g["Ht_bc_ones_f64"]) < 1e-13
    # u is insensitive to the ill-conditioned directions of w (cond ~1e15): 1e-6 is the
    # reproducibility of the reference fp64 solve itself under re-association
    assert o.rel_l2(u, g["u_f64"]) < 1e-6
    assert o.rel_l2(f_pred, g["f_pred_f64"]) < 1e-5


def test_multilambda_vs_reference(golden, weights0):
    g = golden("solve_G32_ml_seed0.npz")
    xp, yp, xb, yb = o.grid_problem(int(g["G"]))
    st = o.precompute(xp, yp, xb, yb, weights0, lambdas=list(g["lambdas"]), dtype=np.float64)
    u = o.solve(st, g["f"], g["bc"])[0]
    assert o.rel_l2(u, g["u_f64"]) < 1e-6


def test_solve_batch_is_column_loop(weights0):
    xp, yp, xb, yb = o.grid_problem(30)
    st = o.precompute(xp, yp, xb, yb, weights0)
    F = np.stack([o.sine_source(xp, yp, seed=s) for s in range(3)], axis=1)
    BC = np.tile(np.array([[1.0, -2.0, 0.5]]), (xb.size, 1))
    U, Fp = o.solve_batch(st, F, BC)
    u1 = o.solve(st, F[:, 1], BC[:, 1])[0]
    assert np.array_equal(U[:, [1]], u1)


def test_range_assert(weights0):
    with pytest.raises(AssertionError, match="x coordinates should be in"):
        o.precompute([0.5, 1.5], [0.5, 0.5], [0.0], [0.0], weights0)


def test_fd_oracle_dst_equals_sparse_lu_restatement():
    """numeric_solve restated with scipy.sparse (numeric_solver.py:104-151) vs the exact DST-I solve."""
    from oracle import fd_oracle as fd

    G = 20
    xp, yp, _, _ = o.grid_problem(G)
    f = o.sine_source(xp, yp, seed=4)
    a = fd.fd_solve_sparse(G, f, -2.5)
    b = fd.fd_solve_dst(G, f, -2.5)
    assert o.rel_l2(a, b) < 1e-12
    F = np.stack([f, 2 * f + 1], 1)
    B = fd.fd_solve_dst(G, F, np.array([-2.5, 4.0]))
    assert o.rel_l2(B[:, 0], a) < 1e-12 and B.shape == (G * G, 2)
    # second-order convergence to the analytic solution of u_xx + u_yy = 50 sin(2 pi x) sin(3 pi y), u = 3 on the boundary
    errs = []
    for g in (16, 32):
        x, y, _, _ = o.grid_problem(g)
        exact = 3 - 50 / (13 * np.pi ** 2) * np.sin(2 * np.pi * x) * np.sin(3 * np.pi * y)
        errs.append(o.rel_l2(fd.fd_solve_dst(g, 50 * np.sin(2 * np.pi * x) * np.sin(3 * np.pi * y), 3.0), exact))
    assert 3.0 < errs[0] / errs[1] < 4.5


@pytest.mark.parametrize("name", ["sources_G32_seed0.npz", "sources_G100_seed0.npz"])
def test_oracle_vs_reference_on_baseline_source_families(golden, weights0, name):
    """BASELINE's own source families (six draws of the reference's sincos 'random' family, four Perlin-family fields)
    through the unmodified reference in fp64 (oracle/make_golden.py sources): the oracle reproduces u to the fp64
    reference's own reproducibility on this cond ~1e15 system, orders of magnitude below the reference's fp32 run."""
    import os
    from conftest import GOLDEN

    if not os.path.exists(os.path.join(GOLDEN, name)):
        pytest.skip(f"{name} not minted")
    g = golden(name)
    xp, yp, xb, yb = o.grid_problem(int(g["G"]))
    st = o.precompute(xp, yp, xb, yb, weights0, dtype=np.float64)
    for j, nm in enumerate(g["names"]):
        u = o.solve(st, g["F"][:, j], np.full_like(xb, g["bc"][j]))[0]
        e = o.rel_l2(u, g["U_f64"][:, [j]])
        floor = o.rel_l2(g["U_f32"][:, [j]], g["U_f64"][:, [j]])
        assert e < 1e-5 and e < 0.05 * floor, (str(nm), e, floor)


def test_fd_oracle_vs_reference_numeric_solve(golden):
    """oracle/fd_oracle.py against the reference's OWN numeric_solve (numeric_solver.py:39-160, run through
    oracle/make_golden.py numeric): sparse-LU restatement and exact DST agree with it to solver rounding."""
    from oracle import fd_oracle as fd

    g = golden("numeric_solve_seed0.npz")
    for G in (48, 200):
        xp, yp, xb, yb = o.grid_problem(G)
        for tag in ("sine", "smooth"):
            f, bcv, u_ref = g[f"f_G{G}_{tag}"], float(g[f"bc_G{G}_{tag}"]), g[f"u_G{G}_{tag}"]
            assert u_ref.shape == (G * G + 4 * G + 4,)
            u = fd.fd_solve_dst(G, f, bcv)
            assert o.rel_l2(u, u_ref[: G * G]) < 1e-10, (G, tag)
            assert np.all(u_ref[G * G:] == bcv)
            if G == 48:
                assert o.rel_l2(fd.fd_solve_sparse(G, f, bcv), u_ref[: G * G]) < 1e-10
    # G = 400 (the README's 5 s case): analytic source, u stored at stride 7
    G = 400
    xp, yp, _, _ = o.grid_problem(G)
    f = 50.0 * np.sin(2 * np.pi * xp) * np.sin(3 * np.pi * yp)
    u = np.concatenate([fd.fd_solve_dst(G, f, 3.0), np.full(4 * G + 4, 3.0)])
    assert o.rel_l2(u[::7], g["u_G400_smooth"]) < 1e-9
    u = np.concatenate([fd.fd_solve_dst(G, o.sine_source(xp, yp, seed=G), -2.5), np.full(4 * G + 4, -2.5)])
    assert o.rel_l2(u[::7], g["u_G400_sine"]) < 1e-9
