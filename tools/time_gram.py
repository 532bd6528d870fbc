import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
s = Solver(use_weights=False)
n = 1 << 20
A = torch.randn(n, 704, device="cuda")
G = torch.zeros(704, 704, dtype=torch.float64, device="cuda")
for _ in range(3):
    _lib.check(s._lib.fps_gram_accum(s._ctx, A.data_ptr(), n, 704, G.data_ptr(), None), "gram")
torch.cuda.synchronize()
