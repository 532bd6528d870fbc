"""Developer tool (GPU): exact-int8 Gram (gram_i8.cu) against an fp64 cuBLAS reference, plus timing."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
FP = 704
s = Solver(use_weights=False)
fn = s._lib.fps_gram_accum

def gram(A):
    G = torch.zeros(FP, FP, dtype=torch.float64, device="cuda")
    _lib.check(fn(s._ctx, A.data_ptr(), A.shape[0], FP, G.data_ptr(), None), "gram")
    torch.cuda.synchronize()
    return G

def report(name, A):
    A0 = A.clone()
    G = gram(A)          # snaps A in place
    print("   snapped elements: %.1f %%, max |dA| / colmax %.2e" % (100.0 * (A != A0).float().mean().item(),
          ((A - A0).abs().max(0).values / A0.abs().max(0).values.clamp_min(1e-30)).max().item()))
    Ad = A.double()
    R = Ad.T @ Ad
    d = torch.sqrt(torch.diag(R)).clamp_min(1e-300)
    err = ((G - R).abs() / (d[:, None] * d[None, :]))[:700, :700]
    sym = (G - G.T).abs().max().item()
    print(f"{name}: n={A.shape[0]} max err/sqrt(GiiGjj) {err.max().item():.3e}  mean {err.mean().item():.3e}  asym {sym:.1e}  "
          f"rel-fro {((G-R)[:700,:700].norm()/R[:700,:700].norm()).item():.3e}")

torch.manual_seed(0)
for n in (16384, 20000, 50001, 131072):
    rng = np.random.default_rng(n)
    x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
    H, DH = s._features(x, y, True)
    report("DH features", DH)
    if n == 50001:
        report("H features", H)
        A = torch.randn(n, FP, device="cuda") * torch.pow(10.0, torch.empty(FP, device="cuda").uniform_(-6, 6))[None, :]
        A = A * torch.exp(3 * torch.randn(n, FP, device="cuda"))     # heavy tails: max/rms in the hundreds
        A[:, 700:] = 0
        report("wide-range synthetic", A.contiguous())
        A2 = torch.zeros(n, FP, device="cuda"); A2[:, :10] = 1.0
        report("mostly zero columns", A2)
n = 1 << 20
A = torch.randn(n, FP, device="cuda")
for _ in range(2): gram(A)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
_lib.check(s._lib.fps_profile_enable(s._ctx, 0), "p")
ev[0].record()
G = torch.zeros(FP, FP, dtype=torch.float64, device="cuda")
for _ in range(5):
    _lib.check(fn(s._ctx, A.data_ptr(), n, FP, G.data_ptr(), None), "gram")
ev[1].record(); torch.cuda.synchronize()
print("gram 2^20 points: %.3f ms per call (FPS_GRAM_I8=%s)" % (ev[0].elapsed_time(ev[1]) / 5, os.environ.get("FPS_GRAM_I8")))
