"""CPU oracle for the fast-poisson-solver hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (numpy, plus torch-CPU for the autograd-faithful variant and the
weight initialiser) of the reference algorithm behind ``Solver.precompute`` / ``Solver.solve``.
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``fast_poisson_solver_b200/`` imports it; the product path has no CPU fallback.

Parity status: the reference ships no golden vectors for this path (every reference
test is a smoke test).  The oracle is therefore pinned against *outputs of the reference itself*,
generated in the build container by ``oracle/make_golden.py`` (which imports the unmodified
reference from /root/reference through an import shim) and committed under ``tests/golden/``;
``tests/test_oracle.py`` re-checks the oracle against those files on every run.

Reference citations (paths relative to /root/reference/fast_poisson_solver/):
  architecture .............. resources/infos.yaml:838-859, model.py:23-67, model.py:113-132
  feature map h(x,y) ........ model.py:74-79 (cat, mapping, Sequential), model.py:134-146 (Fourier)
  Laplacian DH .............. utils.py:54-68 (definition: d2/dx2 + d2/dy2 of every feature)
  Gram / boundary terms ..... solver.py:162-178
  LHS / RHS assembly ........ solver.py:187-199
  solve ..................... solver.py:294-331
  input canonicalisation .... utils.py:97-114
  metrics ................... analyze.py:107-119
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------------
# Architecture of the shipped network (resources/infos.yaml:838-859).  ffm is set, so
# model.py:38-46 gives layers_h = [700]*6 and the Sequential holds depth-1 = 5 Linear(700,700)
# (+Dropout(0)+Tanh) blocks at indices 1,4,7,10,13 (index 0 is nn.Identity, model.py:50-51).
# ---------------------------------------------------------------------------------------------
WIDTH = 700
DEPTH = 6
FFM_NUM = 200
FFM_SCALE = 10.0
OUT = 800
N_HIDDEN = DEPTH - 1
HIDDEN_KEYS = [f"model.{1 + 3 * i}" for i in range(N_HIDDEN)]  # model.1, model.4, ... model.13
DEFAULT_LAMBDA = 2.0 ** -12  # solver.py:74-75


def init_weights(seed: int = 0) -> dict:
    """Default-initialised weights exactly as the reference builds them with ``use_weights=False``.

    ``Solver.__init__`` seeds torch with ``seed`` (solver.py:86-88) and then ``build_model``
    constructs ``PINN`` (solver.py:139).  Parameter creation order inside PINN.__init__
    (model.py:38-64) is: FourierFeatureMapping [freqs = randn(2,200)/2*scale (model.py:124),
    coefficients = ones, biases = zeros (model.py:127-128), linear = Linear(400,700)
    (model.py:130)], then five Linear(700,700) (model.py:53), then out = Linear(700,800)
    (model.py:67).  Re-creating the same torch objects in the same order under the same seed
    reproduces the reference's state_dict bit for bit (checked in tests/test_oracle.py against a
    digest taken from the reference's own PINN).  Returns float32 numpy arrays keyed by the
    reference's state_dict names.
    """
    import torch

    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        freqs = torch.randn(2, FFM_NUM) / 2 * FFM_SCALE
        coefficients = torch.ones(FFM_NUM)
        biases = torch.zeros(FFM_NUM)
        w = {
            "mapping.freqs": freqs,
            "mapping.coefficients": coefficients,
            "mapping.biases": biases,
        }
        lin = torch.nn.Linear(2 * FFM_NUM, WIDTH)
        w["mapping.linear.weight"], w["mapping.linear.bias"] = lin.weight.detach(), lin.bias.detach()
        for key in HIDDEN_KEYS:
            lin = torch.nn.Linear(WIDTH, WIDTH)
            w[key + ".weight"], w[key + ".bias"] = lin.weight.detach(), lin.bias.detach()
        lin = torch.nn.Linear(WIDTH, OUT)  # unused by Solver (model.py:67), kept for key parity
        w["out.weight"], w["out.bias"] = lin.weight.detach(), lin.bias.detach()
    finally:
        torch.random.set_rng_state(gen_state)
    return {k: v.numpy().copy() for k, v in w.items()}


def normalise_state_dict(sd: dict) -> dict:
    """Strip 'module.' / '_orig_mod.' prefixes (solver.py:143-148) and convert to numpy."""
    out = {}
    for k, v in sd.items():
        k = k.replace("module.", "").replace("_orig_mod.", "")
        out[k] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    return out


def format_input(v, dtype):
    """utils.py:97-114: anything array-like -> (n,1) column of ``dtype``."""
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=dtype).reshape(-1, 1)


# ---------------------------------------------------------------------------------------------
# Feature map + Laplacian, forward mode.  Mathematically identical to
# model.py:74-79 followed by utils.py:54-68; validated against the reference's autograd.
# ---------------------------------------------------------------------------------------------
def features_forward(x, y, w: dict, dtype=np.float64, laplace: bool = True):
    """Return (H, DH) with H = model.h(x,y) (N,700) and DH = d2H/dx2 + d2H/dy2 (N,700).

    Streams carried through every layer: value v, v_x, v_y and v_lap = v_xx + v_yy.
      Fourier layer (model.py:136-139):  z = x F0 + y F1 + beta ; phi = [c sin z, c cos z]
      affine (model.py:144, :53):        a = v W^T + b  (bias on the value stream only)
      tanh (model.py:155):               h = tanh a, d1 = 1-h^2, d2 = -2 h d1
                                         h_x = d1 a_x, h_y = d1 a_y,
                                         h_lap = d2 (a_x^2 + a_y^2) + d1 a_lap
    All arithmetic in ``dtype`` (weights are cast, as model.to(precision) does, solver.py:151).
    """
    x = format_input(x, dtype)
    y = format_input(y, dtype)
    F = w["mapping.freqs"].astype(dtype)
    c = w["mapping.coefficients"].astype(dtype)[None, :]
    beta = w["mapping.biases"].astype(dtype)[None, :]
    z = x * F[0][None, :] + y * F[1][None, :] + beta
    s, co = np.sin(z), np.cos(z)
    v = np.concatenate([c * s, c * co], axis=1)
    if laplace:
        f0, f1 = F[0][None, :], F[1][None, :]
        vx = np.concatenate([c * co * f0, -c * s * f0], axis=1)
        vy = np.concatenate([c * co * f1, -c * s * f1], axis=1)
        k2 = np.concatenate([f0 * f0 + f1 * f1] * 2, axis=1)
        vl = -k2 * v
    layers = [("mapping.linear.weight", "mapping.linear.bias")] + [
        (k + ".weight", k + ".bias") for k in HIDDEN_KEYS
    ]
    for wk, bk in layers:
        Wt = w[wk].astype(dtype).T
        b = w[bk].astype(dtype)[None, :]
        h = np.tanh(v @ Wt + b)
        if laplace:
            ax, ay, al = vx @ Wt, vy @ Wt, vl @ Wt
            d1 = 1 - h * h
            d2 = -2 * h * d1
            vx, vy, vl = d1 * ax, d1 * ay, d2 * (ax * ax + ay * ay) + d1 * al
        v = h
    return (v, vl) if laplace else (v, None)


def features_autograd(x, y, w: dict, dtype=np.float64, threads: int | None = None):
    """Algorithm-faithful port of the reference's precompute inner loop, used as the CPU baseline.

    Same structure as solver.py:162-166 + utils.py:54-68: build the graph of h(x,y) with torch
    modules, then for each of the 700 features run four ``autograd.grad`` sweeps (u_x, u_y with
    create_graph, then u_xx, u_yy).  2 800 reverse sweeps - this is what the reference spends
    >99.9 % of precompute in.  Only practical for N <~ 10^4.
    """
    import torch
    from torch.autograd import grad

    if threads:
        torch.set_num_threads(threads)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    tw = {k: torch.from_numpy(np.ascontiguousarray(v)).to(tdt) for k, v in w.items()}
    xt = torch.from_numpy(format_input(x, dtype)).requires_grad_(True)
    yt = torch.from_numpy(format_input(y, dtype)).requires_grad_(True)
    xy = torch.cat((xt, yt), dim=1)  # model.py:75
    z = xy @ tw["mapping.freqs"] + tw["mapping.biases"]  # model.py:136
    phi = torch.cat((tw["mapping.coefficients"] * torch.sin(z), tw["mapping.coefficients"] * torch.cos(z)), dim=1)
    h = torch.tanh(torch.nn.functional.linear(phi, tw["mapping.linear.weight"], tw["mapping.linear.bias"]))
    for k in HIDDEN_KEYS:
        h = torch.tanh(torch.nn.functional.linear(h, tw[k + ".weight"], tw[k + ".bias"]))
    cols = []
    for i in range(h.shape[1]):  # utils.py:56
        u_i = h[:, [i]]
        ones = torch.ones_like(u_i)
        u_x = grad(u_i, xt, create_graph=True, grad_outputs=ones, retain_graph=True)[0]
        u_y = grad(u_i, yt, create_graph=True, grad_outputs=ones, retain_graph=True)[0]
        u_xx = grad(u_x, xt, create_graph=False, grad_outputs=torch.ones_like(u_x), retain_graph=True)[0]
        u_yy = grad(u_y, yt, create_graph=False, grad_outputs=torch.ones_like(u_y), retain_graph=True)[0]
        cols.append(u_xx + u_yy)
    DH = torch.cat(cols, dim=1).detach()
    return h.detach().numpy(), DH.numpy()


# ---------------------------------------------------------------------------------------------
# precompute / solve algebra (solver.py:162-199, 294-331)
# ---------------------------------------------------------------------------------------------
class OracleState:
    """The subset of reference ``Solver`` attributes that the hot path produces."""

    __slots__ = ("H", "DH", "DHtDH", "H_bc", "Ht_bc_ones", "LHSs", "NO", "NdO", "lambdas", "dtype")


def precompute(x_pde, y_pde, x_bc, y_bc, w: dict, lambdas=None, dtype=np.float64, laplace="forward") -> OracleState:
    """solver.py:209-257 without the pickle / warm-up: features, Gram, boundary block, LHS.

    ``RHSs = lambda/N * DH^T`` (solver.py:197) is *not* materialised; ``solve`` applies the scale.
    The centering matrix M = I - 11^T/N_bc (solver.py:190-191) is applied as "subtract the column
    mean of H_bc", which is the same product M @ H_bc without the N_bc x N_bc temporary.
    """
    if lambdas is None:
        lambdas = [DEFAULT_LAMBDA]
    xs = [format_input(v, dtype) for v in (x_pde, y_pde, x_bc, y_bc)]
    x_tot = np.concatenate([xs[0], xs[2]])
    y_tot = np.concatenate([xs[1], xs[3]])
    assert x_tot.min() >= 0 and x_tot.max() <= 1, "x coordinates should be in [0, 1]. Please rescale."  # solver.py:244-245
    assert y_tot.min() >= 0 and y_tot.max() <= 1, "y coordinates should be in [0, 1]. Please rescale."  # solver.py:246-247
    st = OracleState()
    st.dtype = np.dtype(dtype)
    if laplace == "autograd":
        st.H, st.DH = features_autograd(xs[0], xs[1], w, dtype)
    else:
        st.H, st.DH = features_forward(xs[0], xs[1], w, dtype)
    st.DHtDH = st.DH.T @ st.DH  # solver.py:167
    st.H_bc, _ = features_forward(xs[2], xs[3], w, dtype, laplace=False)  # solver.py:173-175
    st.Ht_bc_ones = st.H_bc.sum(axis=0, keepdims=True)  # solver.py:176-178  (1,700)
    st.NO = st.DH.shape[0]
    st.NdO = st.H_bc.shape[0]
    MH = st.H_bc - st.H_bc.mean(axis=0, keepdims=True)  # == M @ H_bc, solver.py:190-196
    bc_block = (st.H_bc.T @ MH) / st.NdO
    st.lambdas = list(lambdas)
    lam = np.asarray(lambdas, dtype=dtype).reshape(-1, 1, 1)
    st.LHSs = lam / st.NO * st.DHtDH[None] + bc_block[None]  # solver.py:195-196
    return st


def solve(st: OracleState, f, bc):
    """solver.py:294-331 for one right-hand side.  Returns (u, u_pde, u_bc, f_pred, w_out, bias).

    With several lambdas the one with the smallest loss is kept (solver.py:309-325).
    """
    dtype = st.dtype
    f = format_input(f, dtype)
    bc = format_input(bc, dtype)
    bct_ones = bc.sum()  # solver.py:302 (sums *all* of bc)
    best = None
    for i, lam in enumerate(st.lambdas):
        rhs = (dtype.type(lam) / st.NO) * (st.DH.T @ f)  # solver.py:197 + :301
        w_out = np.linalg.solve(st.LHSs[i], rhs)  # solver.py:305
        bias = -(st.Ht_bc_ones @ w_out - bct_ones) / st.NdO  # solver.py:306
        if len(st.lambdas) > 1:
            u_bc = st.H_bc @ w_out + bias
            f_pred = st.DH @ w_out
            loss = np.mean((f_pred - f) ** 2) + np.mean((u_bc - bc) ** 2)  # solver.py:312-314
            if best is None or loss < best[0]:
                best = (loss, w_out, bias)
        else:
            best = (0.0, w_out, bias)
    _, w_out, bias = best
    u_pde = st.H @ w_out + bias  # solver.py:327
    f_pred = st.DH @ w_out  # solver.py:329
    u_bc = st.H_bc @ w_out + bias  # solver.py:330
    u = np.concatenate([u_pde, u_bc])  # solver.py:331
    return u, u_pde, u_bc, f_pred, w_out, bias


def solve_batch(st: OracleState, F, BC):
    """Batched extension oracle: loop the single-RHS reference algebra over columns
    (the reference cannot batch: utils.py:110-111 flattens, solver.py:302 sums all of bc)."""
    F = np.asarray(F)
    BC = np.asarray(BC)
    U, Fp = [], []
    for j in range(F.shape[1]):
        u, _, _, fp, _, _ = solve(st, F[:, j], BC[:, j])
        U.append(u)
        Fp.append(fp)
    return np.concatenate(U, axis=1), np.concatenate(Fp, axis=1)


# ---------------------------------------------------------------------------------------------
# Problem generators restating the *layout* of data.py:192-217 and a vectorised sine source in
# the family of data_utils/source_functions.py:24-73 (inputs to the path, not part of it).
# ---------------------------------------------------------------------------------------------
def grid_problem(G: int):
    """Interior G x G grid and the 4G+4 boundary ring laid out as data.py:192-217."""
    g = np.linspace(0.0, 1.0, G + 2)
    xi, yi = np.meshgrid(g[1:-1], g[1:-1])
    x_pde, y_pde = xi.flatten(), yi.flatten()
    yg = g[1:-1]
    x_bc = np.concatenate([g, g, np.zeros_like(yg), np.ones_like(yg)])
    y_bc = np.concatenate([np.zeros_like(g), np.ones_like(g), yg, yg])
    return x_pde, y_pde, x_bc, y_bc


def sine_source(x, y, seed: int = 0, terms: int | None = None):
    """sum_i s_i sin(k_i1 pi (x-o_i1)) {sin|cos}(k_i2 pi (y-o_i2)) with k~U(-10,10), o~U(-1,1),
    s~U(-100,100) (parameter ranges of source_functions.py:24-73)."""
    rng = np.random.default_rng(seed)
    n = terms or int(rng.integers(1, 10))
    f = np.zeros_like(np.asarray(x, dtype=np.float64))
    for _ in range(n):
        k1, k2 = rng.uniform(-10, 10, 2)
        o1, o2 = rng.uniform(-1, 1, 2)
        s = rng.uniform(-100, 100)
        second = np.sin if rng.integers(0, 2) else np.cos
        f = f + s * np.sin(k1 * np.pi * (x - o1)) * second(k2 * np.pi * (y - o2))
    return f


def metrics(pred, true):
    """MSE / RMSE / MAE / rMAE as analyze.py:107-119 with normalize=False."""
    pred = np.asarray(pred, dtype=np.float64).reshape(-1)
    true = np.asarray(true, dtype=np.float64).reshape(-1)
    mse = float(np.mean((true - pred) ** 2))
    mae = float(np.mean(np.abs(true - pred)))
    return {"mse": mse, "rmse": mse ** 0.5, "mae": mae, "rmae": mae / max(float(np.mean(np.abs(true))), 1e-6)}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
