"""Developer tool (GPU): ragged-size sweep of the exact int8 paths (Gram, projection, reconstruction) against fp64."""
import os, sys, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_poisson_solver_b200 import Solver, _lib
os.makedirs("/tmp/fps_cal", exist_ok=True); os.chdir("/tmp/fps_cal")
FP = 704
s = Solver(use_weights=False)
lib, ctx = s._lib, s._ctx
torch.manual_seed(1)
bad = 0
def grid(Fd):
    e = (torch.floor(torch.log2(Fd.abs().max(0).values.clamp_min(1e-300))) + 1).cpu().numpy().astype(int)
    up = torch.tensor(np.ldexp(1.0, 30 - e), dtype=torch.float64, device="cuda")
    dn = torch.tensor(np.ldexp(1.0, e - 30), dtype=torch.float64, device="cuda")
    return torch.round(Fd * up[None]) * dn[None]
for n in (16384, 16385, 16511, 16512, 20001, 28672, 28673, 33333, 65537, 131071):
    A = (torch.randn(n, FP, device="cuda") * torch.exp(1.5 * torch.randn(n, FP, device="cuda"))).contiguous()
    A[:, 700:] = 0
    G = torch.zeros(FP, FP, dtype=torch.float64, device="cuda")
    _lib.check(lib.fps_gram_accum(ctx, A.data_ptr(), n, FP, G.data_ptr(), None), "gram")
    Ad = A.double()
    R = Ad.T @ Ad
    d = torch.sqrt(torch.diag(R)).clamp_min(1e-300)
    eg = ((G - R).abs() / (d[:, None] * d[None, :]))[:700, :700].max().item()
    msgs = [f"gram {eg:.1e}"]
    ok = eg < 1e-13
    for B in (32, 33, 63, 64, 65, 255, 257, 1000):
        if n > 40000 and B not in (33, 257):
            continue
        F = torch.randn(n, B, device="cuda") * torch.exp(torch.randn(n, B, device="cuda"))
        Rp = torch.zeros(FP, B, dtype=torch.float64, device="cuda")
        _lib.check(lib.fps_project_rhs(ctx, A.data_ptr(), n, FP, F.data_ptr(), B, B, Rp.data_ptr(), B, None), "proj")
        Fg = grid(F.double())
        ref = Ad.T @ Fg
        sc = Ad.norm(dim=0)[:, None] * Fg.norm(dim=0)[None, :]
        ep = ((Rp - ref).abs() / sc.clamp_min(1e-300))[:700].max().item()
        W = torch.randn(FP, B, dtype=torch.float64, device="cuda") * torch.pow(10.0, torch.empty(FP, 1, dtype=torch.float64, device="cuda").uniform_(0, 5))
        W[700:] = 0
        bias = torch.randn(B, dtype=torch.float64, device="cuda")
        out = torch.empty(n, B, dtype=torch.float32, device="cuda")
        _lib.check(lib.fps_reconstruct(ctx, A.data_ptr(), n, FP, W.data_ptr(), B, bias.data_ptr(), B, out.data_ptr(), B, None), "recon")
        refo = A.double() @ W + bias[None]
        # W is used on a 39-bit grid relative to max_k |w_kj| 2^(e_k) (e_k = exponent of column k of A): every term
        # moves by at most 2^-39 of that, 704 terms; plus the fp32 rounding of the output
        ek = torch.floor(torch.log2(Ad.abs().max(0).values.clamp_min(1e-300))) + 1
        Bj = (W.abs() * torch.pow(2.0, ek)[:, None]).max(0).values
        er = ((out.double() - refo).abs() - 6.0e-8 * refo.abs() - 704 * 2.0 ** -39 * Bj[None, :]).max().item()
        good = ep < 1e-13 and er <= 0
        ok &= good
        if not good:
            msgs.append(f"B={B}: proj {ep:.1e} recon excess {er:.1e}")
    bad += (not ok)
    print(("OK  " if ok else "BAD ") + f"n={n}: " + "; ".join(msgs), flush=True)
print("sweep done, failures:", bad)
