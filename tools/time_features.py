"""Developer tool (GPU): time only fps_features (per kernel class) under the current FPS_TC_* environment."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944 * 16
rng = np.random.default_rng(0)
x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
s = Solver(use_weights=False, chunk_points=int(os.environ.get("CHUNK", "0")))
for _ in range(2):
    s._features(x, y, True)
torch.cuda.synchronize()
_lib.check(s._lib.fps_profile_enable(s._ctx, 1), "p"); _lib.check(s._lib.fps_profile_reset(s._ctx), "p")
reps = int(os.environ.get("REPS", "3"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Clocks
clk = Clocks(0); clk.start()
import time as _t; _t.sleep(0.3)
for _ in range(reps):
    s._features(x, y, True)
torch.cuda.synchronize()
clkres = clk.stop()
pw = [float(r[2]) for r in clk.rows if len(r) > 2 and r[2].replace('.', '').isdigit()]
print("clocks", clkres, "power_w max %.0f" % (max(pw) if pw else -1))
out = {}
for k, name in enumerate(_lib.PROF_KINDS[:4]):
    ms, cnt = C.c_double(), C.c_int64()
    s._lib.fps_profile_read(s._ctx, k, C.byref(ms), C.byref(cnt))
    out[name] = round(ms.value / reps, 3)
tot = sum(out.values())
print({k: os.environ.get(k) for k in os.environ if k.startswith("FPS_TC")}, out, "total %.2f ms -> %.2f Mpts/s" % (tot, n / tot / 1e3))
if os.environ.get("FPS_TC_STATS") == "1":
    import ctypes
    buf = (ctypes.c_ulonglong * 16)()
    s._lib.fps_debug_tc_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
    s._lib.fps_debug_tc_stats(buf, 0)
    names = ["mma_wait_full", "mma_wait_tempty", "mma_loop_total", "acc_wait_tfull", "acc_promote", "acc_finalize", "tiles", "prod_wait_empty", "mma_wait_full_first_slab"] + ["-"] * 7
    v = list(buf)
    tiles = max(1, v[6])
    print({n: round(x / tiles) for n, x in zip(names, v) if n not in ("tiles", "-")}, "cycles per tile over", v[6], "tiles (CTA 0, all layers)")
    print("all leaders: mean loop cycles per launch %.0f, max %d, n %d" % (v[9] / max(1, v[11]), v[10], v[11]))
