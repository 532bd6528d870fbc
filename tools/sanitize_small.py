"""Developer tool (GPU): a small pass through every kernel family (ragged sizes) for compute-sanitizer memcheck."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
rng = np.random.default_rng(0)
n, nb, B = 16411, 517, 37           # ragged: not multiples of 64 / 128 / 256
x, y = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
t = rng.uniform(0, 1, nb); xb, yb = t, np.zeros(nb)
s = Solver(use_weights=False)
s.precompute(x, y, xb, yb)
F = rng.standard_normal((n, B)).astype(np.float32)
u = s.solve(F, rng.uniform(-1, 1, B))[0]
u1 = s.solve(F[:, 0], np.zeros(nb))[0]
ue = s.evaluate(x[:1000], y[:1000])
torch.cuda.synchronize()
print("sanitize pass ok", float(u.abs().max()), float(u1.abs().max()), tuple(ue.shape))
