"""Developer tool (GPU): single-solve latency breakdown at a given size."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
rng = np.random.default_rng(0)
x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
g = np.linspace(0, 1, 402); xb = torch.tensor(np.concatenate([g, g]), device="cuda"); yb = torch.tensor(np.concatenate([0 * g, 0 * g + 1]), device="cuda")
f = torch.randn(n, device="cuda"); bc = torch.zeros(xb.numel(), device="cuda")
s = Solver(use_weights=False)
s.precompute(x, y, xb, yb)
for _ in range(5):
    s.solve(f, bc)
_lib.check(s._lib.fps_profile_enable(s._ctx, 1), "p"); _lib.check(s._lib.fps_profile_reset(s._ctx), "p")
ts = []
for _ in range(20):
    ts.append(s.solve(f, bc)[4])
out = {}
for k, name in enumerate(_lib.PROF_KINDS):
    ms, cnt = C.c_double(), C.c_int64()
    s._lib.fps_profile_read(s._ctx, k, C.byref(ms), C.byref(cnt))
    if cnt.value:
        out[name] = round(ms.value / 20, 4)
print(f"n={n}: solve runtime median {np.median(ts)*1e3:.3f} ms; kernels {out} sum {sum(out.values()):.3f} ms")
