"""Developer tool (GPU): one batched projection + reconstruction call at configs[2] size (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
s = Solver(use_weights=False)
n, B = 262144, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
A = torch.randn(n, 704, device="cuda"); A[:, 700:] = 0
F = torch.randn(n, B, device="cuda")
R = torch.zeros(704, B, dtype=torch.float64, device="cuda")
W = torch.randn(704, B, dtype=torch.float64, device="cuda"); W[700:] = 0
out = torch.empty(n, B, device="cuda")
for _ in range(2):
    _lib.check(s._lib.fps_project_rhs(s._ctx, A.data_ptr(), n, 704, F.data_ptr(), B, B, R.data_ptr(), B, None), "proj")
    _lib.check(s._lib.fps_reconstruct(s._ctx, A.data_ptr(), n, 704, W.data_ptr(), B, None, B, out.data_ptr(), B, None), "recon")
torch.cuda.synchronize()
