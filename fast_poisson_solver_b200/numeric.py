"""`numeric_solve` on the GPU: the reference's finite-difference accuracy oracle (numeric_solver.py:39-160).

Same call as the reference - ``numeric_solve(f, x_pde, y_pde, u_bc, x_bc, y_bc, precision, verbose)`` returning
``(u, dt)`` with ``u`` over [PDE points ; boundary points] in the caller's order - but the 5-point Dirichlet system the
reference hands to a sparse LU (numeric_solver.py:112, :138-149; operators of diff_matrices.py:22-59) is solved exactly
by type-I sine transforms on the device (csrc/fd_dst.cu, ``fps_fd_solve``).  5 s for a 400 x 400 grid in the reference
(README), milliseconds here; a 2048 x 2048 grid (BASELINE configs[4]) does not fit the reference's LU on a host at all.

What the reference requires is required here too: the points must form a full tensor grid (numeric_solver.py:95-104
takes np.unique of the coordinates) whose outermost rows / columns are the boundary points.  Boundary values need not be
constant: they enter the right-hand side of the neighbouring interior equations exactly as the identity rows of the
reference's system matrix make them do.  Host-side work (sorting, folding the boundary values) is numpy on O(N) data,
like the reference's own lexsort; all O(N^1.5) arithmetic is on the GPU.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import _lib


def fd_solve_grid(solver, Gx, Gy, hx, hy, F, bc=None, batch=None):
    """Interior solution on a Gx x Gy grid for B right-hand sides.  F: device (Gx*Gy, B) float32/float64 tensor, row
    iy*Gx + ix (data.py:192-198 ordering); bc: None or B constant boundary values.  Returns a (Gx*Gy, B) float64 tensor."""
    dev = solver._tdev
    if F.dim() == 1:
        F = F.reshape(-1, 1)
    assert F.shape[0] == Gx * Gy and F.is_cuda and F.stride(1) == 1
    B = F.shape[1]
    U = torch.empty((Gx * Gy, B), dtype=torch.float64, device=dev)
    bcd = None if bc is None else torch.as_tensor(np.asarray(bc, dtype=np.float64).reshape(-1), device=dev)
    if bcd is not None and bcd.numel() != B:
        raise ValueError(f"bc must hold {B} constants")
    if batch is None:
        free, _ = torch.cuda.mem_get_info(solver._dev_index)
        batch = int(max(1, min(B, (free // 4) // (2 * Gx * Gy * 8))))
    nbytes = int(solver._lib.fps_fd_workspace_bytes(Gx, Gy, batch))
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    fn = solver._lib.fps_fd_solve_f32 if F.dtype == torch.float32 else solver._lib.fps_fd_solve
    if F.dtype not in (torch.float32, torch.float64):
        raise TypeError("F must be float32 or float64")
    with torch.cuda.device(solver._dev_index):
        _lib.check(fn(solver._ctx, Gx, Gy, float(hx), float(hy), F.data_ptr(), F.stride(0),
                      None if bcd is None else bcd.data_ptr(), B, U.data_ptr(), B, work.data_ptr(), nbytes,
                      solver._stream()), "fps_fd_solve")
    return U


def numeric_solve(f, x_pde, y_pde, u_bc, x_bc, y_bc, precision=torch.float32, verbose=0, solver=None):
    """Drop-in for fast_poisson_solver.numeric_solve (numeric_solver.py:39-160).  ``solver``: a Solver whose context /
    device is used (one is created with default-initialised weights if omitted - the weights play no role here)."""
    if solver is None:
        from .solver import Solver

        solver = Solver(use_weights=False)
    npd = np.float64
    f, x_pde, y_pde, u_bc, x_bc, y_bc = (
        (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(npd).reshape(-1)
        for v in (f, x_pde, y_pde, u_bc, x_bc, y_bc))
    out_dtype = {torch.float32: np.float32, torch.float64: np.float64}[precision]
    x = np.concatenate([x_pde, x_bc])
    y = np.concatenate([y_pde, y_bc])
    xu, yu = np.unique(x), np.unique(y)                       # numeric_solver.py:95-99
    Nx, Ny = len(xu), len(yu)
    if Nx * Ny != x.size:
        raise ValueError("numeric_solve needs a full tensor grid of points (interior + boundary ring)")
    dx, dy = xu[1] - xu[0], yu[1] - yu[0]                      # numeric_solver.py:101-102
    t0 = time.perf_counter()
    # lexicographic position of every point (numeric_solver.py:120-122 sorts by (y, x))
    ix = np.searchsorted(xu, x)
    iy = np.searchsorted(yu, y)
    full = np.full((Ny, Nx), np.nan)
    rhs = np.concatenate([f, u_bc])                           # numeric_solver.py:128-129
    full[iy, ix] = rhs
    is_b = np.zeros((Ny, Nx), dtype=bool)
    is_b[iy[len(x_pde):], ix[len(x_pde):]] = True
    if not (is_b[0].all() and is_b[-1].all() and is_b[:, 0].all() and is_b[:, -1].all() and not is_b[1:-1, 1:-1].any()):
        raise ValueError("numeric_solve expects the boundary points to be exactly the outer ring of the grid")
    # interior equations with the known boundary neighbours moved to the right-hand side
    Fi = full[1:-1, 1:-1].copy()
    Fi[0, :] -= full[0, 1:-1] / dy ** 2
    Fi[-1, :] -= full[-1, 1:-1] / dy ** 2
    Fi[:, 0] -= full[1:-1, 0] / dx ** 2
    Fi[:, -1] -= full[1:-1, -1] / dx ** 2
    Gx, Gy = Nx - 2, Ny - 2
    Fd = torch.as_tensor(Fi.reshape(-1, 1), dtype=torch.float64, device=solver._tdev)
    Ui = fd_solve_grid(solver, Gx, Gy, dx, dy, Fd, bc=None).reshape(Gy, Gx).cpu().numpy()
    full[1:-1, 1:-1] = Ui
    u = full[iy, ix].astype(out_dtype)                         # back to the caller's order (numeric_solver.py:151)
    dt = time.perf_counter() - t0
    if verbose > 0:
        print(f"Time Numeric Solver: {dt:.6f} s")
    return u, dt
