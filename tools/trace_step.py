"""Developer tool (GPU): where does the step's wall time go BETWEEN the kernels?  One bench step (precompute + single
solve, 2^20 points) under torch.profiler (CUPTI): device-busy time, idle gaps above a threshold and the kernels around
them.  Numbers taken under a profiler are diagnostics, never bench values."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver
from bench import perlin, ring, sobol_points, GRID
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
thr_us = float(sys.argv[2]) if len(sys.argv) > 2 else 15.0
x, y = sobol_points(n, seed=0); xb, yb = ring(GRID)
d = lambda a, dt: torch.as_tensor(a, dtype=dt, device="cuda")
xd, yd, xbd, ybd = d(x, torch.float64), d(y, torch.float64), d(xb, torch.float64), d(yb, torch.float64)
fd, bcd = d(perlin(x, y, seed=0), torch.float32), d(np.full(xb.size, 3.0), torch.float32)
s = Solver(use_weights=False)
def step():
    s.precompute(xd, yd, xbd, ybd)
    return s.solve(fd, bcd)
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy, cur_end, gaps = 0.0, t0, []
for i, e in enumerate(ev):
    st, en = e.time_range.start, e.time_range.end
    if st > cur_end:
        if st - cur_end > thr_us:
            gaps.append((st - cur_end, ev[i - 1].name[:60] if i else "-", e.name[:60]))
        busy += en - st
    else:
        busy += max(0.0, en - cur_end)
    cur_end = max(cur_end, en)
print(f"2 steps: span {(t1 - t0) / 1e3:.2f} ms, device busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms, "
      f"{len(ev)} device activities")
gaps.sort(reverse=True)
print(f"gaps > {thr_us} us: {len(gaps)}, total {sum(g[0] for g in gaps) / 1e3:.2f} ms")
for g in gaps[:40]:
    print(f"  {g[0]:8.1f} us   after {g[1]:60s} before {g[2]}")
