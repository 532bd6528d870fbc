"""Mint golden vectors from the UNMODIFIED reference  --  TEST INFRASTRUCTURE, build container only.

Run here (the build container, where /root/reference exists); the GPU box never runs this:

    python oracle/make_golden.py features          # seconds
    python oracle/make_golden.py solve --grid 32   # ~1-2 min   (fp64 + fp32)
    python oracle/make_golden.py solve --grid 100  # ~15-40 min (fp64 + fp32; BASELINE config 1)
    python oracle/make_golden.py sources --grid 100  # ~35 min: ONE reference precompute per precision, then the
                                                     # reference's solve on BASELINE's own source families
    python oracle/make_golden.py numeric           # ~1 min: the reference's finite-difference numeric_solve

What it does: copies /root/reference/fast_poisson_solver to a temp dir
(never into the repo), registers empty stub modules for the plotting/metrics imports the package
pulls in at import time (matplotlib, skimage, sklearn, perlin_noise - all absent here), writes a
placeholder ``resources/final.pt`` (the real one is missing from the checkout and
``Solver.__init__`` opens it unconditionally, solver.py:97-99), then drives the reference's own
``Solver(device='cpu', use_weights=False, compile_model=False, seed=0)`` so that the weights are
the reference's default initialisation under seed 0 - reproducible on any box from
``oracle.fps_oracle.init_weights(0)`` without the reference.  Outputs go to tests/golden/*.npz.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import sys
import tempfile
import time
import types

import numpy as np

REF = "/root/reference/fast_poisson_solver"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    """Return the reference package imported from a throw-away copy (with import stubs)."""
    import torch  # noqa: F401

    tmp = tempfile.mkdtemp(prefix="fps_ref_")
    shutil.copytree(REF, os.path.join(tmp, "fast_poisson_solver"))

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    stub("matplotlib", rcParams={})
    stub("matplotlib.pyplot")
    stub("matplotlib.font_manager")
    stub("matplotlib.cm", ScalarMappable=object)
    stub("skimage")
    stub("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)
    stub("sklearn")
    stub("sklearn.metrics", r2_score=None)
    stub("perlin_noise", PerlinNoise=object)
    sys.path.insert(0, tmp)
    cwd = os.getcwd()
    os.chdir(tmp)  # Solver() creates ./precomputed-resources (solver.py:105-107)
    try:
        import fast_poisson_solver as ref  # noqa
        from fast_poisson_solver.model import PINN  # noqa
        import yaml

        # placeholder weight file so that Solver() can open it; never loaded (use_weights=False)
        with open(os.path.join(tmp, "fast_poisson_solver", "resources", "infos.yaml")) as fh:
            yaml.SafeLoader.add_constructor(
                "tag:yaml.org,2002:python/tuple", lambda loader, node: tuple(loader.construct_sequence(node))
            )
            infos = yaml.safe_load(fh)
        torch.save(PINN(infos["model"]).state_dict(), os.path.join(tmp, "fast_poisson_solver", "resources", "final.pt"))
    finally:
        os.chdir(cwd)
    return ref, tmp


def digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def new_solver(ref, tmp, precision, lambdas=None):
    import torch

    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        return ref.Solver(device="cpu", precision=precision, use_weights=False, compile_model=False,
                          lambdas_pde=lambdas, seed=0)
    finally:
        os.chdir(cwd)


def cmd_features(ref, tmp):
    """(i) H, DH for 48 random points through model.h + calculate_laplace, fp64 and fp32,
    plus the sha256 of the reference's default-init state_dict under seed 0."""
    import torch
    from fast_poisson_solver.utils import calculate_laplace

    rng = np.random.default_rng(7)
    x = rng.uniform(0, 1, 48)
    y = rng.uniform(0, 1, 48)
    out = {"x": x, "y": y}
    for name, prec in (("f64", torch.float64), ("f32", torch.float32)):
        s = new_solver(ref, tmp, prec)
        if name == "f64":
            sd = {k: v.to(torch.float32) for k, v in s.model.state_dict().items()}
            out["weights_sha256"] = np.array(digest(sd))
        xt = torch.tensor(x, dtype=prec).reshape(-1, 1).requires_grad_(True)
        yt = torch.tensor(y, dtype=prec).reshape(-1, 1).requires_grad_(True)
        H = s.model.h(xt, yt)
        DH = calculate_laplace(H, xt, yt).detach()
        out[f"H_{name}"] = H.detach().numpy()
        out[f"DH_{name}"] = DH.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "features_seed0.npz"), **out)
    print("wrote features_seed0.npz", out["weights_sha256"])


def cmd_solve(ref, tmp, G, lambdas=None, tag=""):
    """(ii) full reference precompute + solve on a G x G grid (data.py:192-217 layout),
    f = 50 sin(2 pi x) sin(3 pi y), u_bc = 3, fp64 (parity target) and
    fp32 (noise floor).  Stores u, f_pred, Ht_bc_ones, a strided sample of DHtDH and LHS, timings."""
    import torch

    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.fps_oracle import grid_problem

    x_pde, y_pde, x_bc, y_bc = grid_problem(G)
    f = 50.0 * np.sin(2 * np.pi * x_pde) * np.sin(3 * np.pi * y_pde)
    bc = np.full_like(x_bc, 3.0)
    out = {"G": np.array(G), "f": f, "bc": bc, "threads": np.array(torch.get_num_threads())}
    if lambdas is not None:
        out["lambdas"] = np.array(lambdas)
    for name, prec in (("f64", torch.float64), ("f32", torch.float32)):
        s = new_solver(ref, tmp, prec, lambdas)
        t0 = time.perf_counter()
        s.precompute(x_pde.copy(), y_pde.copy(), x_bc.copy(), y_bc.copy())
        t1 = time.perf_counter()
        u, u_pde, u_bc, f_pred, rt = s.solve(f.copy(), bc.copy())
        out[f"u_{name}"] = u.detach().numpy()
        out[f"f_pred_{name}"] = f_pred.detach().numpy()
        out[f"Ht_bc_ones_{name}"] = s.Ht_bc_ones.detach().numpy()
        out[f"DHtDH_s7_{name}"] = s.DHtDH.detach().numpy()[::7, ::7]
        out[f"LHS_s7_{name}"] = s.LHSs[0].detach().numpy()[::7, ::7]
        out[f"H_rows_{name}"] = s.H.detach().numpy()[:: max(1, G * G // 16)][:16]
        out[f"DH_rows_{name}"] = s.DH.detach().numpy()[:: max(1, G * G // 16)][:16]
        out[f"bias_{name}"] = np.asarray(float(np.asarray(s.bias).reshape(-1)[0]))
        out[f"lambda_{name}"] = np.asarray(float(s.lambda_pde))
        out[f"t_precompute_{name}"] = np.array(t1 - t0)
        out[f"t_solve_{name}"] = np.array(rt)
        print(f"G={G} {name}: precompute {t1 - t0:.1f}s solve {rt * 1e3:.2f} ms", flush=True)
    np.savez_compressed(os.path.join(GOLDEN, f"solve_G{G}{tag}_seed0.npz"), **out)
    print(f"wrote solve_G{G}{tag}_seed0.npz")


def cmd_sources(ref, tmp, G):
    """(iv) BASELINE's own source families through the unmodified reference, fp64 and fp32, on a G x G grid:
    ONE reference precompute per precision (the expensive 2 800-sweep part does not depend on f), then
    Solver.solve for
      * six draws of the reference's own sine family, source_functions.sincos(x, y, 'random') under
        np.random.seed(0) (k ~ U(-10,10), 1-9 terms, source_functions.py:24-73), boundary value ~ U(-10,10) as
        data.py:224-225;
      * four Perlin-family fields (source_functions.py:173-227 parameter draws; lattice gradients from this repo's
        hash-based generator because the perlin_noise package is absent - sources are inputs);
      * the benign 50 sin(2 pi x) sin(3 pi y) of the other goldens as a control.
    Stored per source: f, bc, u_f64, u_f32 (+ f_pred_f64).  tests assert  err(ours, ref64) <= max(1e-5, err(ref32, ref64))."""
    import torch

    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.fps_oracle import grid_problem
    from fast_poisson_solver.data_utils.source_functions import sincos
    from fast_poisson_solver_b200.sources import perlin_params, perlin_sources_numpy

    x_pde, y_pde, x_bc, y_bc = grid_problem(G)
    np.random.seed(0)
    F, BC, names = [], [], []
    for i in range(6):
        bcv = np.random.uniform(-10, 10)                      # data.py:224-225
        F.append(sincos(x_pde.copy(), y_pde.copy(), "random"))
        BC.append(bcv)
        names.append(f"sine{i}")
    pp = perlin_params(4, seed=11)
    Fp = perlin_sources_numpy(x_pde, y_pde, pp)
    for i in range(4):
        F.append(Fp[:, i])
        BC.append(np.random.uniform(-10, 10))
        names.append(f"perlin{i}")
    F.append(50.0 * np.sin(2 * np.pi * x_pde) * np.sin(3 * np.pi * y_pde))
    BC.append(3.0)
    names.append("benign")
    F = np.stack(F, 1)
    BC = np.asarray(BC)
    out = {"G": np.array(G), "F": F, "bc": BC, "names": np.array(names), "perlin_params": pp}
    for name, prec in (("f64", torch.float64), ("f32", torch.float32)):
        s = new_solver(ref, tmp, prec)
        t0 = time.perf_counter()
        s.precompute(x_pde.copy(), y_pde.copy(), x_bc.copy(), y_bc.copy())
        print(f"G={G} {name}: precompute {time.perf_counter() - t0:.1f}s", flush=True)
        U, FP, WN = [], [], []
        for j in range(F.shape[1]):
            u, u_pde, u_bc, f_pred, rt = s.solve(F[:, j].copy(), np.full_like(x_bc, BC[j]))
            U.append(u.detach().numpy().reshape(-1))
            FP.append(f_pred.detach().numpy().reshape(-1))
            WN.append(float(s.w_out.abs().max()))
        out[f"U_{name}"] = np.stack(U, 1)
        out[f"Fpred_{name}"] = np.stack(FP, 1)
        out[f"wmax_{name}"] = np.asarray(WN)
    for j, nm in enumerate(names):
        a, b = out["U_f32"][:, j].astype(np.float64), out["U_f64"][:, j]
        print(f"  {nm}: reference fp32 vs fp64 u rel-L2 = {np.linalg.norm(a - b) / np.linalg.norm(b):.3e}, |w|max {out['wmax_f64'][j]:.3g}")
    np.savez_compressed(os.path.join(GOLDEN, f"sources_G{G}_seed0.npz"), **out)
    print(f"wrote sources_G{G}_seed0.npz")


def cmd_numeric(ref, tmp):
    """(v) the reference's OWN finite-difference solver, numeric_solve (numeric_solver.py:39-160), in fp64 on grids
    of data.py:192-217: pins oracle/fd_oracle.py and the GPU DST solver (csrc/fd_dst.cu).  G = 48 and 200 with stored
    sources; G = 400 (the README's 5 s case) with an analytic source, u stored at stride 7."""
    import torch

    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.fps_oracle import grid_problem, sine_source
    from fast_poisson_solver.numeric_solver import numeric_solve

    out = {}
    for G, stride in ((48, 1), (200, 1), (400, 7)):
        x_pde, y_pde, x_bc, y_bc = grid_problem(G)
        for tag, f, bcv in (("sine", sine_source(x_pde, y_pde, seed=G), -2.5),
                            ("smooth", 50.0 * np.sin(2 * np.pi * x_pde) * np.sin(3 * np.pi * y_pde), 3.0)):
            t0 = time.perf_counter()
            u, dt = numeric_solve(f.copy(), x_pde.copy(), y_pde.copy(), np.full_like(x_bc, bcv), x_bc.copy(), y_bc.copy(),
                                  precision=torch.float64)
            u = np.asarray(u).reshape(-1)
            print(f"numeric_solve G={G} {tag}: {time.perf_counter() - t0:.2f}s (reported {dt:.2f}s)", flush=True)
            out[f"u_G{G}_{tag}"] = u[::stride]
            out[f"bc_G{G}_{tag}"] = np.asarray(bcv)
            if stride == 1:
                out[f"f_G{G}_{tag}"] = f
        out[f"stride_G{G}"] = np.asarray(stride)
    np.savez_compressed(os.path.join(GOLDEN, "numeric_solve_seed0.npz"), **out)
    print("wrote numeric_solve_seed0.npz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["features", "solve", "multilambda", "sources", "numeric"])
    ap.add_argument("--grid", type=int, default=32)
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    ref, tmp = import_reference()
    try:
        if a.what == "features":
            cmd_features(ref, tmp)
        elif a.what == "solve":
            cmd_solve(ref, tmp, a.grid)
        elif a.what == "sources":
            cmd_sources(ref, tmp, a.grid)
        elif a.what == "numeric":
            cmd_numeric(ref, tmp)
        else:
            cmd_solve(ref, tmp, a.grid, lambdas=[2.0 ** -12, 2.0 ** -8, 2.0 ** -4], tag="_ml")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
