"""Developer tool (GPU): exact-int8 batched RHS projection (gram_i8.cu) against an fp64 cuBLAS reference, plus timing."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
FP = 704
s = Solver(use_weights=False)
def project(DH, F):
    B = F.shape[1]
    R = torch.zeros(FP, B, dtype=torch.float64, device="cuda")
    _lib.check(s._lib.fps_project_rhs(s._ctx, DH.data_ptr(), DH.shape[0], FP, F.data_ptr(), B, B, R.data_ptr(), B, None), "proj")
    torch.cuda.synchronize()
    return R
for n, B in ((16384, 32), (20001, 100), (65536, 512), (50000, 4096)):
    rng = np.random.default_rng(n)
    x = torch.tensor(rng.uniform(0, 1, n), device="cuda"); y = torch.tensor(rng.uniform(0, 1, n), device="cuda")
    H, DH = s._features(x, y, True)
    G = torch.zeros(FP, FP, dtype=torch.float64, device="cuda")
    _lib.check(s._lib.fps_gram_accum(s._ctx, DH.data_ptr(), n, FP, G.data_ptr(), None), "gram")
    DH1 = DH.clone()
    F = (torch.randn(n, B, device="cuda") * torch.exp(torch.randn(n, B, device="cuda"))).contiguous()
    F0 = F.clone()
    R = project(DH, F)
    assert (DH == DH1).all() and (F == F0).all()
    ref = DH.double().T @ F.double()
    sc = DH.double().norm(dim=0)[:, None] * F.double().norm(dim=0)[None, :]
    err = ((R - ref).abs() / sc.clamp_min(1e-300))[:700]
    print(f"n={n} B={B}: max |R-ref|/(|dh_f||f_b|) {err.max().item():.3e}  rel-fro {((R-ref)[:700].norm()/ref[:700].norm()).item():.3e}")
n, B = 262144, 4096
DH = torch.randn(n, FP, device="cuda"); F = torch.randn(n, B, device="cuda")
for _ in range(2): project(DH, F)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(3): project(DH, F)
ev[1].record(); torch.cuda.synchronize()
print("project 262144 x 4096: %.3f ms per call" % (ev[0].elapsed_time(ev[1]) / 3))
