import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fast_poisson_solver_b200 import Solver
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
for n in (1 << 20, 1 << 21, 1 << 22):
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
    g = np.linspace(0, 1, 2048)
    xb = torch.tensor(np.concatenate([g, g]), device="cuda"); yb = torch.tensor(np.concatenate([0 * g, 0 * g + 1]), device="cuda")
    s = Solver(use_weights=False)
    s.precompute(x, y, xb, yb)
    f = torch.randn(n, device="cuda"); bc = torch.full((xb.numel(),), 2.5, device="cuda")
    s.solve(f, bc)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = s.solve(f, bc); e1.record(); torch.cuda.synchronize()
        ts.append((round(e0.elapsed_time(e1), 3), round(out[4] * 1e3, 3)))
    print(n, ts, flush=True)
    del s, x, y, f
    torch.cuda.empty_cache()
