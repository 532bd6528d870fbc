"""Finite-difference accuracy oracle  --  TEST INFRASTRUCTURE ONLY (same rules as fps_oracle.py).

Restates the reference's `numeric_solve` (numeric_solver.py:39-160) for the grids of data.py:192-217:
the system matrix is the 5-point Laplacian D2x/dx^2 + D2y/dy^2 (numeric_solver.py:112, interior
rows of diff_matrices.py:31, :47-53) on the full (G+2) x (G+2) tensor grid with EVERY boundary row
replaced by an identity row (numeric_solver.py:138-146), right-hand side = [f ; u_bc] in
lexicographic order (numeric_solver.py:120-131), solved with a sparse LU (numeric_solver.py:149).

`fd_solve_sparse` is that algorithm with scipy.sparse; `fd_solve_dst` solves the SAME linear system
exactly with a type-I discrete sine transform (the interior block is the standard Dirichlet 5-point
operator, whose eigenvectors are sines), which is what makes the 2048^2 check of BASELINE config 5
feasible.  tests/test_oracle.py checks the two against each other.
"""
from __future__ import annotations

import numpy as np


def _grid(G):
    g = np.linspace(0.0, 1.0, G + 2)
    return g, g[1] - g[0]


def fd_solve_sparse(G: int, f_interior, bc_value: float):
    """u on the interior G x G points (row-major like data.py:192-198: y slow, x fast)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import spsolve

    n = G + 2
    _, h = _grid(G)
    d2 = sp.diags([1.0, -2.0, 1.0], [-1, 0, 1], shape=(n, n))
    L = (sp.kron(sp.eye(n), d2) + sp.kron(d2, sp.eye(n))) / h ** 2
    L = L.tolil()
    idx = np.arange(n * n).reshape(n, n)
    boundary = np.concatenate([idx[0], idx[-1], idx[1:-1, 0], idx[1:-1, -1]])
    rhs = np.zeros(n * n)
    rhs[idx[1:-1, 1:-1].reshape(-1)] = np.asarray(f_interior, dtype=np.float64).reshape(-1)
    for b in boundary:
        L.rows[b] = [int(b)]
        L.data[b] = [1.0]
    rhs[boundary] = bc_value
    u = spsolve(L.tocsr(), rhs)
    return u.reshape(n, n)[1:-1, 1:-1].reshape(-1)


def fd_solve_dst(G: int, f_interior, bc_value):
    """Same system, solved exactly by diagonalising the Dirichlet 5-point operator with DST-I.

    f_interior: (G*G,) or (G*G, B); bc_value: scalar or (B,) constants.  With a constant boundary
    value c, u = c + v where v solves the zero-Dirichlet problem (the stencil annihilates constants).
    """
    from scipy.fft import dstn, idstn

    _, h = _grid(G)
    F = np.asarray(f_interior, dtype=np.float64)
    single = F.ndim == 1
    F = F.reshape(G, G, -1)
    k = np.arange(1, G + 1)
    lam = (2.0 * np.cos(np.pi * k / (G + 1)) - 2.0) / h ** 2
    denom = lam[:, None] + lam[None, :]
    V = idstn(dstn(F, type=1, axes=(0, 1)) / denom[:, :, None], type=1, axes=(0, 1))
    U = V + np.asarray(bc_value, dtype=np.float64).reshape(1, 1, -1)
    U = U.reshape(G * G, -1)
    return U[:, 0] if single else U
