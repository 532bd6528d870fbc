"""Developer tool (GPU): exact-int8 batched reconstruction (recon_i8.cu) against an fp64 reference, plus timing."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
FP = 704
s = Solver(use_weights=False)
def recon(A, W, bias):
    n, B = A.shape[0], W.shape[1]
    out = torch.empty(n, B, dtype=torch.float32, device="cuda")
    _lib.check(s._lib.fps_reconstruct(s._ctx, A.data_ptr(), n, FP, W.data_ptr(), B, bias.data_ptr() if bias is not None else None, B,
                                      out.data_ptr(), B, None), "recon")
    torch.cuda.synchronize()
    return out
torch.manual_seed(0)
for n, B in ((16384, 32), (20001, 100), (65536, 512), (30000, 1000)):
    rng = np.random.default_rng(n)
    x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
    H, DH = s._features(x, y, True)
    for name, A in (("H", H), ("DH", DH)):
        # weights with heavy cancellation: large entries, output much smaller than the terms
        W = torch.randn(FP, B, dtype=torch.float64, device="cuda") * torch.pow(10.0, torch.empty(FP, 1, device="cuda", dtype=torch.float64).uniform_(0, 5))
        W[700:] = 0
        bias = torch.randn(B, dtype=torch.float64, device="cuda")
        out = recon(A, W, bias)            # snaps A in place
        ref = A.double() @ W + bias[None]
        terms = (A.double().abs() @ W.abs())   # size of the summands
        err = (out.double() - ref).abs()
        print(f"{name} n={n} B={B}: max err / sum|terms| {(err / terms).max().item():.2e}; max err / |ref| rms {(err.max() / ref.pow(2).mean().sqrt()).item():.2e}; "
              f"fp32 rounding floor {6e-8:.0e}")
n, B = 262144, 4096
A = torch.randn(n, FP, device="cuda"); W = torch.randn(FP, B, dtype=torch.float64, device="cuda"); bias = torch.zeros(B, dtype=torch.float64, device="cuda")
for _ in range(2): recon(A, W, bias)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(3): recon(A, W, bias)
ev[1].record(); torch.cuda.synchronize()
print("reconstruct 262144 x 4096: %.3f ms per call" % (ev[0].elapsed_time(ev[1]) / 3))
