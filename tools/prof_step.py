"""Developer tool (GPU): one short pass over every kernel of the path, for ncu captures (tools/profile_r02.sh).
4 waves of collocation points (151 552), precompute twice, one single solve, two batched solves of 512 right-hand sides."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rng = np.random.default_rng(0)
x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
g = np.linspace(0, 1, 402)
xb = torch.tensor(np.concatenate([g, g]), device="cuda"); yb = torch.tensor(np.concatenate([0 * g, 0 * g + 1]), device="cuda")
s = Solver(use_weights=False)
for _ in range(2):
    s.precompute(x, y, xb, yb)
f = torch.randn(n, device="cuda"); bc = torch.zeros(xb.numel(), device="cuda")
s.solve(f, bc)
F = torch.randn(n, B, device="cuda"); bcv = torch.zeros(B, device="cuda")
for _ in range(2):
    s.solve(F, bcv)
torch.cuda.synchronize()
print("prof_step done", n, B)
