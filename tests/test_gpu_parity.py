"""GPU: parity of the CUDA path (through the C ABI / the Solver) against the CPU oracle and the
golden vectors minted from the reference.  Tolerances: north_star asks u within 1e-5 relative L2 of
the reference (fp64 reference run = the well-defined arithmetic)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fps_oracle as o

pytestmark = pytest.mark.gpu

ENGINES = ["simt", "tcgen05"]
TOL_FEATURE = 4e-6   # fp32-grade features vs fp64 oracle (reference fp32 itself sits at 5e-7 / 9e-7; 3xTF32 ~1e-6 / 3e-6)
TOL_U = 1e-5         # north_star parity bar


@pytest.fixture(scope="module")
def solver_factory():
    from fast_poisson_solver_b200 import Solver

    made = []

    def make(**kw):
        kw.setdefault("use_weights", False)
        s = Solver(**kw)
        made.append(s)
        return s

    yield make
    made.clear()


@pytest.mark.parametrize("engine", ENGINES)
def test_features_vs_reference_golden(golden, solver_factory, engine):
    g = golden("features_seed0.npz")
    s = solver_factory(engine=engine)
    x = torch.tensor(g["x"], device="cuda")
    y = torch.tensor(g["y"], device="cuda")
    H, DH = s._features(x, y, True)
    Hb, none = s._features(x, y, False)
    torch.cuda.synchronize()
    assert none is None
    H, DH, Hb = H.cpu().numpy(), DH.cpu().numpy(), Hb.cpu().numpy()
    assert np.all(H[:, 700:] == 0) and np.all(DH[:, 700:] == 0) and np.all(Hb[:, 700:] == 0)
    assert o.rel_l2(H[:, :700], g["H_f64"]) < TOL_FEATURE
    assert o.rel_l2(DH[:, :700], g["DH_f64"]) < TOL_FEATURE
    assert o.rel_l2(Hb[:, :700], g["H_f64"]) < TOL_FEATURE


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n", [1, 127, 129, 1000, 20000])
def test_features_ragged_sizes_vs_oracle(weights0, solver_factory, engine, n):
    """Tile tails, multi-wave chunks (chunk_points=4096 -> 5 waves at n=20000)."""
    rng = np.random.default_rng(n)
    x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    s = solver_factory(engine=engine, chunk_points=4096)
    H, DH = s._features(torch.tensor(x, device="cuda"), torch.tensor(y, device="cuda"), True)
    Hr, DHr = o.features_forward(x, y, weights0, np.float64)
    assert o.rel_l2(H[:, :700].cpu().numpy(), Hr) < TOL_FEATURE
    assert o.rel_l2(DH[:, :700].cpu().numpy(), DHr) < TOL_FEATURE
    # per-point error, so that a single bad tile cannot hide in the norm
    err = np.linalg.norm(DH[:, :700].cpu().numpy() - DHr, axis=1) / np.linalg.norm(DHr, axis=1)
    assert err.max() < 2e-5


def test_engines_agree_on_large_batch(solver_factory):
    n = 50000
    rng = np.random.default_rng(5)
    x = torch.tensor(rng.uniform(0, 1, n), device="cuda")
    y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
    a = solver_factory(engine="simt")._features(x, y, True)
    b = solver_factory(engine="tcgen05")._features(x, y, True)
    for u, v in zip(a, b):
        assert o.rel_l2(u.cpu().numpy(), v.cpu().numpy()) < 5e-6


def test_gram_colsum_project_reconstruct_kernels(solver_factory):
    """The fp64-accumulating kernels against numpy fp64 on the same fp32 inputs: exact to rounding."""
    from fast_poisson_solver_b200 import _lib

    s = solver_factory(engine="simt")
    lib, ctx, st = s._lib, s._ctx, s._stream()
    rng = np.random.default_rng(0)
    for n in (1, 33, 700, 5000):
        A = torch.zeros((n, 704), dtype=torch.float32, device="cuda")
        A[:, :700] = torch.tensor(rng.standard_normal((n, 700)), dtype=torch.float32)
        G = torch.zeros((704, 704), dtype=torch.float64, device="cuda")
        sv = torch.zeros(704, dtype=torch.float64, device="cuda")
        _lib.check(lib.fps_gram_accum(ctx, A.data_ptr(), n, 704, G.data_ptr(), st), "gram")
        _lib.check(lib.fps_colsum_accum(ctx, A.data_ptr(), n, 704, sv.data_ptr(), st), "colsum")
        A64 = A.double().cpu().numpy()
        assert np.allclose(G.cpu().numpy(), A64.T @ A64, rtol=1e-13, atol=1e-12)
        assert np.allclose(sv.cpu().numpy(), A64.sum(0), rtol=1e-13, atol=1e-12)
        for B in (1, 3, 8, 9, 64, 130):
            Fm = torch.tensor(rng.standard_normal((n, B)), dtype=torch.float32, device="cuda")
            R = torch.zeros((704, B), dtype=torch.float64, device="cuda")
            _lib.check(lib.fps_project_rhs(ctx, A.data_ptr(), n, 704, Fm.data_ptr(), B, B, R.data_ptr(), B, st), "proj")
            assert np.allclose(R.cpu().numpy(), A64.T @ Fm.double().cpu().numpy(), rtol=1e-13, atol=1e-11), (n, B)
            W = torch.tensor(rng.standard_normal((704, B)), dtype=torch.float64, device="cuda")
            bias = torch.tensor(rng.standard_normal(B), dtype=torch.float64, device="cuda")
            out = torch.empty((n, B), dtype=torch.float32, device="cuda")
            _lib.check(lib.fps_reconstruct(ctx, A.data_ptr(), n, 704, W.data_ptr(), B, bias.data_ptr(), B,
                                           out.data_ptr(), B, st), "recon")
            ref = A64 @ W.cpu().numpy() + bias.cpu().numpy()[None]
            assert np.allclose(out.cpu().numpy(), ref.astype(np.float32), rtol=2e-7, atol=1e-5), (n, B)


def test_gram_int8_exact_path(solver_factory):
    """n >= 16384 rows: the Gram runs on the int8 tensor cores (csrc/gram_i8.cu).  It snaps A in place to a
    per-column 31-bit grid and must return the Gram of the stored A to fp64 rounding, reproducibly."""
    from fast_poisson_solver_b200 import _lib

    s = solver_factory()
    lib, ctx, st = s._lib, s._ctx, s._stream()
    rng = np.random.default_rng(5)
    for n in (16384, 40001):
        a = rng.standard_normal((n, 704)) * np.exp(2.0 * rng.standard_normal((n, 704)))   # heavy tails
        a *= 10.0 ** rng.uniform(-5, 5, size=(1, 704))                                     # column scales
        a[:, 700:] = 0.0
        a[:, 17] = 0.0                                                                     # an all-zero column
        A = torch.tensor(a, dtype=torch.float32, device="cuda")
        A0 = A.clone()
        G = torch.zeros((704, 704), dtype=torch.float64, device="cuda")
        _lib.check(lib.fps_gram_accum(ctx, A.data_ptr(), n, 704, G.data_ptr(), st), "gram")
        torch.cuda.synchronize()
        cmax = A0.abs().max(0).values
        assert ((A - A0).abs() <= cmax[None, :] * 2.0 ** -30).all()       # |dA| <= half a grid step of 2^(e-30), e <= log2(cmax)+1
        changed = (A != A0)
        assert (A0.abs()[changed] < (cmax[None, :] * 2.0 ** -6).expand_as(A0)[changed]).all()   # only small elements move
        A64 = A.double().cpu().numpy()
        ref = A64.T @ A64
        d = np.sqrt(np.maximum(np.diag(ref), 1e-300))
        err = np.abs(G.cpu().numpy() - ref) / (d[:, None] * d[None, :])
        assert err[:700, :700].max() < 1e-13, (n, err.max())
        assert (G == G.T).all()
        # second call: A is already on the grid -> unchanged, and the result is bit-identical
        A1 = A.clone()
        G2 = torch.zeros_like(G)
        _lib.check(lib.fps_gram_accum(ctx, A.data_ptr(), n, 704, G2.data_ptr(), st), "gram")
        torch.cuda.synchronize()
        assert (A == A1).all() and (G2 == G).all()


def test_project_and_reconstruct_int8_exact_paths(solver_factory):
    """n >= 16384 and B >= 32: batched projection DH^T F and reconstruction A W + b run on the int8 tensor cores
    (csrc/gram_i8.cu, csrc/recon_i8.cu) with exact integer products.  Checked against fp64 on the operands the kernels
    define: DH / H snapped to their 31-bit column grids (in place), F on a 31-bit grid per right-hand side."""
    from fast_poisson_solver_b200 import _lib

    s = solver_factory()
    lib, ctx, st = s._lib, s._ctx, s._stream()
    rng = np.random.default_rng(9)
    for n, B in ((16384, 32), (16385, 33), (20001, 100), (28673, 257), (33000, 700)):   # incl. ragged tile / slab / segment edges
        x = torch.tensor(rng.uniform(0, 1, n), device="cuda")
        y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
        H, DH = s._features(x, y, True)
        G = torch.zeros((704, 704), dtype=torch.float64, device="cuda")
        _lib.check(lib.fps_gram_accum(ctx, DH.data_ptr(), n, 704, G.data_ptr(), st), "gram")   # snaps DH
        DH1 = DH.clone()
        F = torch.tensor(rng.standard_normal((n, B)) * np.exp(rng.standard_normal((n, B))), dtype=torch.float32, device="cuda")
        F0 = F.clone()
        R = torch.zeros((704, B), dtype=torch.float64, device="cuda")
        _lib.check(lib.fps_project_rhs(ctx, DH.data_ptr(), n, 704, F.data_ptr(), B, B, R.data_ptr(), B, st), "proj")
        torch.cuda.synchronize()
        assert (DH == DH1).all() and (F == F0).all()          # DH already on its grid, F never written
        Fd = F.double()
        e = (torch.floor(torch.log2(Fd.abs().max(0).values)) + 1).cpu().numpy().astype(int)
        up = torch.tensor(np.ldexp(1.0, 30 - e), dtype=torch.float64, device="cuda")      # exact powers of two (a device
        dn = torch.tensor(np.ldexp(1.0, e - 30), dtype=torch.float64, device="cuda")      # pow() is not: it flips ties)
        Fg = torch.round(Fd * up[None, :]) * dn[None, :]
        ref = DH.double().T @ Fg
        scale = DH.double().norm(dim=0)[:, None] * Fg.norm(dim=0)[None, :]
        assert ((R - ref).abs() / scale.clamp_min(1e-300))[:700].max().item() < 1e-13, (n, B)
        assert ((R - DH.double().T @ Fd)[:700].norm() / ref[:700].norm()).item() < 1e-7   # vs the unrounded F
        # reconstruction with heavily cancelling weights
        W = torch.tensor(rng.standard_normal((704, B)) * 10.0 ** rng.uniform(0, 5, size=(704, 1)), dtype=torch.float64, device="cuda")
        W[700:] = 0
        bias = torch.tensor(rng.standard_normal(B), dtype=torch.float64, device="cuda")
        for A in (H, DH):
            out = torch.empty((n, B), dtype=torch.float32, device="cuda")
            _lib.check(lib.fps_reconstruct(ctx, A.data_ptr(), n, 704, W.data_ptr(), B, bias.data_ptr(), B, out.data_ptr(), B, st), "recon")
            torch.cuda.synchronize()
            Ad = A.double()
            ref = Ad @ W + bias[None]                         # A after the call = the snapped operand
            # W is used on a 39-bit grid relative to max_k |w_kj| 2^(e_k), e_k = exponent of column k of A: each of the
            # 704 terms moves by at most 2^-39 of that; the output is rounded to fp32
            ek = torch.floor(torch.log2(Ad.abs().max(0).values.clamp_min(1e-300))) + 1
            Bj = (W.abs() * torch.pow(2.0, ek)[:, None]).max(0).values
            err = (out.double() - ref).abs()
            assert (err <= 6.0e-8 * ref.abs() + 704 * 2.0 ** -39 * Bj[None, :] + 1e-30).all(), (n, B, err.max().item())
        assert (DH == DH1).all()


def test_gram_int8_multipass_in_subprocess():
    """The chunk loop of the int8 Gram (n beyond one pass of digit planes), forced with FPS_GRAM_I8_CHUNK in a fresh process."""
    import subprocess, sys, textwrap

    code = textwrap.dedent("""
        import os, sys, numpy as np, torch
        sys.path.insert(0, %r)
        os.chdir(%r)
        from fast_poisson_solver_b200 import Solver, _lib
        s = Solver(use_weights=False)
        n = 40001
        A = torch.randn(n, 704, device="cuda"); A[:, 700:] = 0
        G = torch.zeros(704, 704, dtype=torch.float64, device="cuda")
        _lib.check(s._lib.fps_gram_accum(s._ctx, A.data_ptr(), n, 704, G.data_ptr(), None), "gram")
        torch.cuda.synchronize()
        R = A.double().T @ A.double()
        err = ((G - R).abs().max() / R.abs().max()).item()
        assert err < 1e-13, err
        print("MULTIPASS_OK", err)
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.getcwd())
    env = dict(os.environ, FPS_GRAM_I8_CHUNK="8192")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MULTIPASS_OK" in r.stdout, r.stdout + r.stderr


def test_cholesky_and_batched_solve_kernels(solver_factory):
    import ctypes as C
    from fast_poisson_solver_b200 import _lib

    s = solver_factory(engine="simt")
    lib, ctx, st = s._lib, s._ctx, s._stream()
    rng = np.random.default_rng(1)
    n, nbc = 3000, 200
    A = rng.standard_normal((n, 700))
    Hb = rng.standard_normal((nbc, 700))
    G = np.zeros((704, 704)); G[:700, :700] = A.T @ A
    Gb = np.zeros((704, 704)); Gb[:700, :700] = Hb.T @ Hb
    sv = np.zeros(704); sv[:700] = Hb.sum(0)
    lam = [0.5, 2.0]
    dG, dGb, dsv = (torch.tensor(v, device="cuda") for v in (G, Gb, sv))
    L = torch.empty((2, _lib.LROWS, 704), dtype=torch.float64, device="cuda")
    info = torch.ones(2, dtype=torch.int32, device="cuda")
    _lib.check(lib.fps_assemble_factor(ctx, dG.data_ptr(), dGb.data_ptr(), dsv.data_ptr(), n, nbc,
                                       (C.c_double * 2)(*lam), None, 2, L.data_ptr(), info.data_ptr(), st), "factor")
    assert info.cpu().tolist() == [0, 0]
    for i, l in enumerate(lam):
        LHS = l / n * G[:700, :700] + (Gb[:700, :700] - np.outer(sv[:700], sv[:700]) / nbc) / nbc
        Lr = np.linalg.cholesky(LHS)
        Ld = L[i, :704].cpu().numpy()
        assert np.allclose(np.tril(Ld[:700, :700]), Lr, rtol=1e-10, atol=1e-12)
        assert np.allclose(Ld, Ld.T)  # mirrored
        assert np.allclose(Ld[700:, 700:], np.eye(4))
        for B in (1, 5, 8, 19):
            R = np.zeros((704, B)); R[:700] = rng.standard_normal((700, B))
            sb = rng.standard_normal(B)
            dR, dsb = torch.tensor(R, device="cuda"), torch.tensor(sb, device="cuda")
            W = torch.empty((704, B), dtype=torch.float64, device="cuda")
            bias = torch.empty(B, dtype=torch.float64, device="cuda")
            _lib.check(lib.fps_solve_factored(ctx, L[i].data_ptr(), dR.data_ptr(), B, B, 0.25, dsv.data_ptr(),
                                              dsb.data_ptr(), nbc, W.data_ptr(), bias.data_ptr(), st), "solve")
            Wr = np.linalg.solve(LHS, 0.25 * R[:700])
            assert np.allclose(W.cpu().numpy()[:700], Wr, rtol=1e-8, atol=1e-10)
            assert np.allclose(bias.cpu().numpy(), -(sv[:700] @ Wr - sb) / nbc, rtol=1e-8, atol=1e-10)
    # a singular matrix must be reported, not produce NaNs
    _lib.check(lib.fps_assemble_factor(ctx, torch.zeros_like(dG).data_ptr(), dGb.data_ptr(), dsv.data_ptr(), n, nbc,
                                       (C.c_double * 1)(1.0), None, 1, L.data_ptr(), info.data_ptr(), st), "factor")
    assert info.cpu().tolist()[0] > 0 and torch.isfinite(L[0]).all()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", ["solve_G32_seed0.npz", "solve_G100_seed0.npz"])
def test_solver_vs_reference_golden(golden, solver_factory, engine, name):
    """BASELINE config 1 shape: unit-square grid, sin source, constant Dirichlet BC; drop-in API."""
    if not os.path.exists(os.path.join(GOLDEN, name)):
        pytest.skip(f"{name} not minted")
    g = golden(name)
    xp, yp, xb, yb = o.grid_problem(int(g["G"]))
    s = solver_factory(engine=engine)
    s.precompute(xp, yp, xb, yb)
    u, u_pde, u_bc, f_pred, rt = s.solve(g["f"], g["bc"])
    N, Nbc = xp.size, xb.size
    assert u.shape == (N + Nbc, 1) and u_pde.shape == (N, 1) and u_bc.shape == (Nbc, 1) and f_pred.shape == (N, 1)
    assert isinstance(rt, float) and u.dtype == torch.float32 and u.is_cuda
    assert s.NO == N and s.NdO == Nbc and s.H.shape == (N, 700) and s.w_out.shape == (700, 1)
    e_u = o.rel_l2(u.cpu().numpy(), g["u_f64"])
    e_f = o.rel_l2(f_pred.cpu().numpy(), g["f_pred_f64"])
    floor = o.rel_l2(g["u_f32"], g["u_f64"])
    print(f"\n[{name} {engine}] u rel-L2 vs reference fp64 = {e_u:.3e} (reference fp32 vs fp64 = {floor:.3e}); f_pred {e_f:.3e}; jitter {s.jitter}")
    assert e_u < TOL_U
    assert e_f < 1e-4
    assert o.rel_l2(s.DHtDH.cpu().numpy()[::7, ::7], g["DHtDH_s7_f64"]) < 1e-5
    assert o.rel_l2(s.Ht_bc_ones.cpu().numpy(), g["Ht_bc_ones_f64"]) < 2e-6
    assert o.rel_l2(s.LHSs[0].cpu().numpy()[::7, ::7], g["LHS_s7_f64"]) < 1e-5


def test_multilambda_vs_reference_golden(golden, solver_factory):
    g = golden("solve_G32_ml_seed0.npz")
    xp, yp, xb, yb = o.grid_problem(int(g["G"]))
    s = solver_factory(engine="simt", lambdas_pde=list(g["lambdas"]))
    s.precompute(xp, yp, xb, yb)
    u = s.solve(g["f"], g["bc"])[0]
    assert s.lambda_pde == float(g["lambda_f64"])
    assert o.rel_l2(u.cpu().numpy(), g["u_f64"]) < TOL_U
    assert np.all(s.Ls > 0)


def _oracle_state_from_device_features(s):
    """Oracle algebra (solver.py:167-199) in fp64 on the features the GPU produced: isolates the
    Gram / factor / projection / reconstruction kernels from feature rounding."""
    st = o.OracleState()
    st.dtype = np.dtype(np.float64)
    st.H, st.DH, st.H_bc = (t.double().cpu().numpy() for t in (s.H, s.DH, s.H_bc))
    st.DHtDH = st.DH.T @ st.DH
    st.Ht_bc_ones = st.H_bc.sum(0, keepdims=True)
    st.NO, st.NdO = st.DH.shape[0], st.H_bc.shape[0]
    MH = st.H_bc - st.H_bc.mean(0, keepdims=True)
    st.lambdas = list(s.lambdas_pde)
    lam = np.asarray(st.lambdas).reshape(-1, 1, 1)
    st.LHSs = lam / st.NO * st.DHtDH[None] + (st.H_bc.T @ MH / st.NdO)[None]
    return st


@pytest.mark.parametrize("engine", ENGINES)
def test_batched_solve_vs_oracle_column_loop(weights0, solver_factory, engine):
    """Batched RHS (extension): oracle = the reference algebra looped over columns.

    Sources are the BASELINE sine family (k ~ U(-10,10)); with default-init weights they drive
    |w| to ~1e5, so u amplifies feature rounding by ~1e3: fp32-grade features (ours at ~6e-7, the
    reference's own fp32 features at 5e-7/9e-7) put u 2e-4..6e-4 away from the all-fp64 reference
    (measured with the CPU oracle, DESIGN.md "precision budget").  So the 1e-5 bar is asserted on
    the algebra with identical features, and the end-to-end number is asserted against the
    documented fp32-feature floor.
    """
    G = 40
    xp, yp, xb, yb = o.grid_problem(G)
    B = 11
    Fm = np.stack([o.sine_source(xp, yp, seed=j) for j in range(B)], axis=1)
    bcv = np.random.default_rng(0).uniform(-10, 10, B)
    BC = np.tile(bcv[None], (xb.size, 1))
    s = solver_factory(engine=engine)
    s.precompute(xp, yp, xb, yb)
    u, u_pde, u_bc, f_pred, _ = s.solve(Fm, BC)
    assert u.shape == (xp.size + xb.size, B) and f_pred.shape == (xp.size, B)
    U, Fp = o.solve_batch(_oracle_state_from_device_features(s), Fm, BC)
    for j in range(B):
        assert o.rel_l2(u[:, j].cpu().numpy(), U[:, j]) < TOL_U, j
        assert o.rel_l2(f_pred[:, j].cpu().numpy(), Fp[:, j]) < 1e-4, j  # f_pred = DH w cancels ~1e3:1
    U64, _ = o.solve_batch(o.precompute(xp, yp, xb, yb, weights0), Fm, BC)
    worst = max(o.rel_l2(u[:, j].cpu().numpy(), U64[:, j]) for j in range(B))
    print(f"\n[batched {engine}] worst u rel-L2 vs all-fp64 reference algebra: {worst:.3e}")
    assert worst < 3e-3
    u2 = s.solve(Fm, bcv)[0]  # constants form
    assert torch.equal(u, u2)
    u1 = s.solve(Fm[:, 3], BC[:, 3])[0]
    assert o.rel_l2(u1.cpu().numpy(), u[:, [3]].cpu().numpy()) < 1e-6


def test_input_forms_and_errors(solver_factory):
    xp, yp, xb, yb = o.grid_problem(28)
    s = solver_factory(engine="simt")
    with pytest.raises(AssertionError, match="x coordinates should be in"):
        s.precompute(xp + 0.5, yp, xb, yb)
    with pytest.raises(AssertionError, match="y coordinates should be in"):
        s.precompute(xp, yp - 0.5, xb, yb)
    f = np.sin(xp * 3) * 10
    s.precompute(list(xp), torch.tensor(yp), xb.reshape(2, -1), torch.tensor(yb, dtype=torch.float32))
    a = s.solve(list(f), list(np.full(xb.size, 2.0)))[0]
    b = s.solve(torch.tensor(f).reshape(28, 28), torch.full((xb.size, 1), 2.0))[0]
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        s.solve(f[:-1], np.zeros(xb.size))


def test_rank_deficient_fixture_shape_does_not_crash(solver_factory):
    """The reference's own test fixture (grid_num=5: 49 equations, 700 unknowns) is singular; the
    reference returns an arbitrary finite LU answer.  We must return finite numbers too."""
    xp, yp, xb, yb = o.grid_problem(5)
    s = solver_factory(engine="simt")
    s.precompute(xp, yp, xb, yb)
    u, _, u_bc, f_pred, _ = s.solve(np.ones(25), np.full(xb.size, 1.5))
    assert torch.isfinite(u).all() and torch.isfinite(f_pred).all()
    assert any(s.jitter)
    assert o.rel_l2(f_pred.cpu().numpy(), np.ones(25)) < 1e-3
    assert abs(u_bc.mean().item() - 1.5) < 1e-3


def test_save_load_roundtrip(solver_factory):
    xp, yp, xb, yb = o.grid_problem(30)
    f = o.sine_source(xp, yp, seed=3)
    s = solver_factory(engine="simt")
    s.precompute(xp, yp, xb, yb, name="rt_case", save=True)
    assert os.path.isfile(s.precompute_file)
    u0 = s.solve(f, np.zeros(xb.size))[0]
    s2 = solver_factory(engine="simt")
    s2.precompute(xp, yp, xb, yb, name="rt_case", load=True)
    assert torch.equal(s2.solve(f, np.zeros(xb.size))[0], u0)
    os.remove(s.precompute_file)
    s2.precompute(xp, yp, xb, yb, name="rt_case", load=True)  # missing file -> silently recompute
    assert torch.equal(s2.solve(f, np.zeros(xb.size))[0], u0)


# ---- precision = torch.float64: full-fp64 engine, no fp32 floor ---------------------------------
def test_f64_features_vs_reference_golden(golden, solver_factory):
    g = golden("features_seed0.npz")
    s = solver_factory(precision=torch.float64)
    H, DH = s._features(torch.tensor(g["x"], device="cuda"), torch.tensor(g["y"], device="cuda"), True)
    Hb, _ = s._features(torch.tensor(g["x"], device="cuda"), torch.tensor(g["y"], device="cuda"), False)
    assert H.dtype == torch.float64 and torch.all(H[:, 700:] == 0) and torch.all(DH[:, 700:] == 0)
    assert o.rel_l2(H[:, :700].cpu().numpy(), g["H_f64"]) < 1e-13
    assert o.rel_l2(DH[:, :700].cpu().numpy(), g["DH_f64"]) < 1e-13
    assert o.rel_l2(Hb[:, :700].cpu().numpy(), g["H_f64"]) < 1e-13


@pytest.mark.parametrize("name", ["solve_G32_seed0.npz", "solve_G100_seed0.npz"])
def test_f64_solver_vs_reference_golden(golden, solver_factory, name):
    g = golden(name)
    xp, yp, xb, yb = o.grid_problem(int(g["G"]))
    s = solver_factory(precision=torch.float64)
    s.precompute(xp, yp, xb, yb)
    u, u_pde, u_bc, f_pred, rt = s.solve(g["f"], g["bc"])
    assert u.dtype == torch.float64 and s.H.dtype == torch.float64 and s.precision == torch.float64
    e_u = o.rel_l2(u.cpu().numpy(), g["u_f64"])
    print(f"\n[{name} fp64 engine] u rel-L2 vs reference fp64 = {e_u:.3e}; jitter {s.jitter}")
    assert e_u < 1e-6      # reproducibility of the reference's own fp64 LU on this cond~1e15 system
    assert o.rel_l2(s.DHtDH.cpu().numpy()[::7, ::7], g["DHtDH_s7_f64"]) < 1e-12
    assert o.rel_l2(f_pred.cpu().numpy(), g["f_pred_f64"]) < 1e-5


def test_f64_rough_sources_meet_1e5_end_to_end(weights0, solver_factory):
    """The BASELINE sine family (|w| ~ 1e5) that no fp32 pipeline can bring under 1e-5: the fp64
    engine does, end to end, against the all-fp64 reference algebra."""
    G, B = 40, 9
    xp, yp, xb, yb = o.grid_problem(G)
    Fm = np.stack([o.sine_source(xp, yp, seed=j) for j in range(B)], axis=1)
    bcv = np.random.default_rng(0).uniform(-10, 10, B)
    BC = np.tile(bcv[None], (xb.size, 1))
    s = solver_factory(precision=torch.float64)
    s.precompute(xp, yp, xb, yb)
    u = s.solve(Fm, BC)[0]
    U64, _ = o.solve_batch(o.precompute(xp, yp, xb, yb, weights0), Fm, BC)
    worst = max(o.rel_l2(u[:, j].cpu().numpy(), U64[:, j]) for j in range(B))
    print(f"\n[fp64 engine, rough sources] worst u rel-L2 vs reference algebra: {worst:.3e}")
    assert worst < TOL_U
    u1 = s.solve(Fm[:, 2], BC[:, 2])[0]
    assert o.rel_l2(u1.cpu().numpy(), U64[:, [2]]) < TOL_U


def test_device_sine_sources_and_evaluate(solver_factory):
    from fast_poisson_solver_b200.sources import sine_params, sine_sources, sine_sources_numpy

    xp, yp, xb, yb = o.grid_problem(36)
    s = solver_factory()
    s.precompute(xp, yp, xb, yb)
    par = sine_params(37, seed=3)
    Fd = sine_sources(s, par)
    Fh = sine_sources_numpy(xp, yp, par)
    assert Fd.shape == (xp.size, 37) and Fd.dtype == torch.float32
    assert np.allclose(Fd.cpu().numpy(), Fh, rtol=1e-6, atol=1e-4)
    u, u_pde, _, _, _ = s.solve(Fd[:, :3], np.array([1.0, -1.0, 0.5]))
    # u(x, y) evaluated at the collocation points reproduces u_pde; elsewhere it is finite and smooth
    again = s.evaluate(xp, yp)
    assert o.rel_l2(again.cpu().numpy(), u_pde.cpu().numpy()) < 1e-6
    mid = s.evaluate(np.array([0.5, 0.25]), np.array([0.5, 0.75]))
    assert mid.shape == (2, 3) and torch.isfinite(mid).all()


def test_fp16_operand_scales_are_calibrated_and_robust(golden, solver_factory):
    """The tensor-core engine stores layer inputs as scaled fp16 hi/lo planes; precompute calibrates the
    scales on the actual points.  A weight set with 30x larger derivative streams (scaled Fourier
    frequencies) must neither overflow nor lose accuracy."""
    import ctypes as C
    from fast_poisson_solver_b200 import weights as W

    g = golden("features_seed0.npz")
    sd = W.default_init(0)
    s = solver_factory(weights=sd)
    xp, yp, xb, yb = o.grid_problem(24)
    s.precompute(xp, yp, xb, yb)
    sc = (C.c_float * 24)()
    s._lib.fps_get_scales(s._ctx, sc)
    sc = np.array(list(sc)).reshape(6, 4)
    assert np.all(np.log2(sc) == np.round(np.log2(sc)))          # powers of two
    assert np.all(sc[:, 3] < sc[:, 0])                            # laplacian stream is the largest
    # stress: 3x the Fourier frequencies -> ~3x / 9x larger d/dx, laplacian streams
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["mapping.freqs"] = sd["mapping.freqs"] * 3.0
    s2 = solver_factory(weights=sd2)
    s2.precompute(xp, yp, xb, yb)
    w2 = {k: v.numpy() for k, v in sd2.items()}
    Hr, DHr = o.features_forward(xp, yp, w2, np.float64)
    assert torch.isfinite(s2.DH).all()
    assert o.rel_l2(s2.H.cpu().numpy(), Hr) < TOL_FEATURE
    assert o.rel_l2(s2.DH.cpu().numpy(), DHr) < TOL_FEATURE


def test_full_size_properties_config2(solver_factory):
    """BASELINE configs[1] size (2^20 scattered points): size-independent properties instead of an oracle run.
    linearity of solve in (f, bc), batch == single, evaluate() == u_pde at the collocation points,
    tensor-core engine == fp32 cross-check engine on a sub-sample, residual of the fitted source."""
    from scipy.stats import qmc

    n = 1 << 20
    pts = qmc.Sobol(d=2, scramble=True, seed=0).random(n)
    x = torch.tensor(pts[:, 0], device="cuda")
    y = torch.tensor(pts[:, 1], device="cuda")
    g = np.linspace(0, 1, 1026)
    xb = np.concatenate([g, g, np.zeros(1024), np.ones(1024)])
    yb = np.concatenate([np.zeros(1026), np.ones(1026), g[1:-1], g[1:-1]])
    s = solver_factory()
    s.precompute(x, y, xb, yb)
    assert s.NO == n and s.NdO == 4100 and s.H.shape == (n, 700) and s.jitter == [0.0]
    f1 = 30.0 * torch.sin(2 * np.pi * x) * torch.sin(3 * np.pi * y)
    f2 = 20.0 * torch.cos(np.pi * x) * y
    F = torch.stack([f1, f2, 2.0 * f1 - 0.5 * f2], dim=1).float()
    bcv = np.array([1.0, -2.0, 2.0 * 1.0 - 0.5 * -2.0])
    u, u_pde, u_bc, f_pred, _ = s.solve(F, bcv)
    assert u.shape == (n + 4100, 3) and torch.isfinite(u).all()
    lin = 2.0 * u[:, 0] - 0.5 * u[:, 1]
    assert (lin - u[:, 2]).norm() / u[:, 2].norm() < 2e-5           # linearity (fp32 outputs)
    sub = torch.arange(0, n, 997, device="cuda")
    ue = s.evaluate(x[sub], y[sub])                                   # uses the weights of the last solve
    assert ue.shape == (sub.numel(), 3)
    assert (ue - u_pde[sub]).norm() / u_pde[sub].norm() < 1e-5       # u(x, y) re-evaluated
    u1 = s.solve(F[:, 1], np.full(4100, -2.0))[0]
    assert (u1[:, 0] - u[:, 1]).norm() / u[:, 1].norm() < 1e-6       # batched == single
    assert (f_pred[:, 0] - F[:, 0]).norm() / F[:, 0].norm() < 1e-2   # the fitted source
    assert (u_bc[:, 0] - 1.0).abs().max() < 1e-2
    ref = solver_factory(engine="simt")
    Hs, DHs = ref._features(x[sub].double(), y[sub].double(), True)
    assert o.rel_l2(s.DH[sub].cpu().numpy(), DHs[:, :700].cpu().numpy()) < 3e-6
    assert o.rel_l2(s.H[sub].cpu().numpy(), Hs[:, :700].cpu().numpy()) < 3e-6


def test_constructor_variants_like_reference_test_solverclass(solver_factory):
    """tests/test_solverclass.py:46-88 of the reference: ctor variants and the attributes it asserts."""
    from fast_poisson_solver_b200 import Solver

    s = solver_factory()
    assert isinstance(s, Solver) and s.device == "cuda" and s.precision == torch.float32
    assert s.lambdas_pde == [2 ** -12] and s.n_lambdas == 1 and os.path.isdir(s.precompute_path)
    s = solver_factory(device=torch.device("cuda"))
    assert s.device == torch.device("cuda")
    s = solver_factory(device="cuda:0", precision=torch.float64, verbose=True, compile_model=False, seed=3)
    assert s.precision == torch.float64 and s.device == "cuda:0" and s.compile_model is False
    with pytest.raises(RuntimeError, match="only CUDA"):
        Solver(device="cpu", use_weights=False)
    with pytest.raises(FileNotFoundError):
        Solver()  # use_weights=True and no final.pt shipped (reference: solver.py:97-99)
    with pytest.raises(RuntimeError, match="precompute"):
        solver_factory().solve([0.0], [0.0])
