"""Developer tool (GPU): throughput of the fp64 engine vs the fp32 engine at a given size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
rng = np.random.default_rng(0)
x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
g = np.linspace(0, 1, 1026); xb = torch.tensor(np.concatenate([g, g]), device="cuda"); yb = torch.tensor(np.concatenate([0 * g, 0 * g + 1]), device="cuda")
f = torch.randn(n, device="cuda"); bc = torch.zeros(xb.numel(), device="cuda")
for prec in (torch.float32, torch.float64):
    s = Solver(use_weights=False, precision=prec)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.precompute(x, y, xb, yb)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        rt = s.solve(f, bc)[4]
    print(f"{prec}: n={n} precompute {1e3*(t1-t0):.1f} ms = {n/(t1-t0)/1e6:.3f} Mpoints/s, single solve {rt*1e3:.3f} ms")
