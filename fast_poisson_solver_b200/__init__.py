"""B200-native (sm_100a) implementation of the fast-poisson-solver hot path.

    from fast_poisson_solver_b200 import Solver
    solver = Solver()                      # same kwargs as the reference's Solver (solver.py:67-68)
    solver.precompute(x_pde, y_pde, x_bc, y_bc)
    u, u_pde, u_bc, f_pred, t = solver.solve(f, u_bc)

Everything numerical runs in libfps_sm100.so (hand-written CUDA, C ABI in include/fps_sm100.h).
"""
from .solver import Solver, format_input  # noqa: F401
from .numeric import numeric_solve  # noqa: F401  (GPU restatement of the reference's FD accuracy oracle)

__all__ = ["Solver", "format_input", "numeric_solve"]
__version__ = "0.2.0"
