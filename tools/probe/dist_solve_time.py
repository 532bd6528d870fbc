"""torchrun probe: per-call times of a distributed single solve (points sharded), every rank prints its own."""
import os, sys, time, datetime
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
from fast_poisson_solver_b200 import Solver
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
rng = np.random.default_rng(rank)
x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
g = np.linspace(0, 1, 2048)
xb = torch.tensor(np.concatenate([g, g]), device="cuda"); yb = torch.tensor(np.concatenate([0 * g, 0 * g + 1]), device="cuda")
s = Solver(device=f"cuda:{local}", use_weights=False, distributed=True)
s.precompute(x, y, xb, yb)
f = torch.randn(n, device="cuda"); bc = torch.full((xb.numel(),), 2.5, device="cuda")
s.solve(f, bc)
rows = []
for _ in range(8):
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); out = s.solve(f, bc); e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    rows.append((round(e0.elapsed_time(e1), 2), round(out[4] * 1e3, 2), round((t1 - t0) * 1e3, 2)))
print(f"rank {rank} n {n}: (event ms, solve() runtime ms, wall ms) {rows}", flush=True)
# the same with the profile of the library switched on (per kernel class)
import ctypes as C
from fast_poisson_solver_b200 import _lib
_lib.check(s._lib.fps_profile_enable(s._ctx, 1), "p"); _lib.check(s._lib.fps_profile_reset(s._ctx), "p")
for _ in range(4):
    s.solve(f, bc)
torch.cuda.synchronize()
out = {}
for k, name in enumerate(_lib.PROF_KINDS):
    ms, cnt = C.c_double(), C.c_int64()
    s._lib.fps_profile_read(s._ctx, k, C.byref(ms), C.byref(cnt))
    if cnt.value: out[name] = round(ms.value / 4, 3)
print(f"rank {rank} kernels per solve {out}", flush=True)
dist.barrier()
dist.destroy_process_group()
