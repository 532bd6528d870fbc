"""Developer tool (GPU): time fps_features per kernel class AND measure H / DH against the full-fp64 engine on a subset,
under the current FPS_* environment (FPS_TC_SEG, FPS_K0_LD, ...).  One line of JSON per run.

    FPS_TC_SEG=4 python tools/exp_features.py [points] [check_points]
"""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 37888 * 8
nchk = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
rng = np.random.default_rng(0)
x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
g = np.linspace(0, 1, 402)
xb = torch.tensor(np.concatenate([g, g]), device="cuda"); yb = torch.tensor(np.concatenate([0 * g, 0 * g + 1]), device="cuda")
s = Solver(use_weights=False, chunk_points=int(os.environ.get("CHUNK", "0")))
for _ in range(2):
    s.precompute(x, y, xb, yb)           # the production path: fused digit split in the last layer's epilogue
torch.cuda.synchronize()
_lib.check(s._lib.fps_profile_enable(s._ctx, 1), "p"); _lib.check(s._lib.fps_profile_reset(s._ctx), "p")
reps = int(os.environ.get("REPS", "3"))
for _ in range(reps):
    s.precompute(x, y, xb, yb)
torch.cuda.synchronize()
H, DH = s._Hp, s._DHp
out = {}
for k, name in enumerate(_lib.PROF_KINDS[:6]):
    ms, cnt = C.c_double(), C.c_int64()
    s._lib.fps_profile_read(s._ctx, k, C.byref(ms), C.byref(cnt))
    out[name] = round(ms.value / reps * (1 << 20) / n, 3)          # ms per 2^20 points
out["features"] = round(sum(out[k] for k in ("fourier", "layer0", "hidden", "last")), 3)
_lib.check(s._lib.fps_profile_enable(s._ctx, 0), "p")
# accuracy against the fp64 engine on the first nchk points
s64 = Solver(use_weights=False, precision=torch.float64)
H64, DH64 = s64._features(x[:nchk].contiguous(), y[:nchk].contiguous(), True)
rel = lambda a, b: float((a.double() - b).norm() / b.norm())
out["H_rel"] = rel(H[:nchk, :700], H64[:, :700]); out["DH_rel"] = rel(DH[:nchk, :700], DH64[:, :700])
cal = (C.c_double * 5)(); s._lib.fps_get_calibration(s._ctx, cal)
out["cal"] = [float(v) for v in cal]
out["env"] = {k: v for k, v in os.environ.items() if k.startswith("FPS_")}
if os.environ.get("FPS_TC_STATS") == "1":
    buf = (C.c_ulonglong * 16)()
    s._lib.fps_debug_tc_stats.argtypes = [C.c_void_p, C.c_int]
    s._lib.fps_debug_tc_stats(buf, 0)
    names = ["mma_wait_full", "mma_wait_tempty", "mma_loop_total", "acc_wait_tfull", "acc_promote", "acc_finalize", "tiles",
             "prod_wait_empty", "mma_wait_full_first_slab"]
    v = list(buf); tiles = max(1, v[6])
    out["stats_per_tile"] = {nm: round(val / tiles) for nm, val in zip(names, v) if nm != "tiles"}
print(json.dumps(out), flush=True)
