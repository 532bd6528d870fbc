"""Drop-in ``Solver`` for the fast-poisson-solver hot path on B200 (sm_100a).

Public surface mirrors the reference class (paths relative to /root/reference/fast_poisson_solver/):
  constructor kwargs and defaults ........ solver.py:67-68, lambdas default [2**-12] :74-75
  precompute(x_pde,y_pde,x_bc,y_bc,...) .. solver.py:209-265 (range asserts :244-247)
  solve(f, bc) -> 5-tuple ................ solver.py:267-337
  attributes callers read ................ H, DH, H_bc, DHtDH, Ht_bc_ones, NO, NdO, LHSs, RHSs,
                                           w_out, bias, lambda_pde, Ls, Ls_pde, Ls_bc, u_pred, ...

All arithmetic of the path runs in hand-written CUDA kernels behind the C ABI of
``libfps_sm100.so`` (include/fps_sm100.h); torch is used for device memory, streams and
``torch.distributed`` plumbing only.  There is no CPU fallback: constructing a Solver without a
CUDA device raises.

Deliberate deviations from the reference:
  * CUDA only; ``compile_model`` is accepted and ignored (the reference's torch.compile never
    touches ``model.h`` anyway).
  * ``solve`` also accepts ``f`` of shape (N, B) with ``bc`` of shape (N_bc, B) (or (B,) constants)
    and returns (N+N_bc, B) / (N, B) / (N_bc, B) tensors - the reference is single-RHS only.
  * ``runtime`` is measured with a stream synchronise on both sides.
  * precision=torch.float32: features are evaluated at fp32 grade (tensor-core contractions on fp16 hi/lo
    operand planes, three MMAs per product; fp64 Fourier phases); Gram matrix, RHS projection and
    reconstruction are fp64 grade - exact integer arithmetic on the int8 tensor cores for N >= 16384 points
    (and >= 32 right-hand sides), fp64 accumulation otherwise - and the factorisation is fp64.
    precision=torch.float64: everything, including the feature network, runs in fp64 (DMMA).
  * For N >= 16384 the stored ``DH`` (and ``H`` once a batched solve has used it) is snapped to a 31-bit
    fixed-point grid per feature column: elements below 2^-7 of the column maximum move by less than 2^-31 of
    it (under 1 % of one fp32 ulp of the maximum), all others are unchanged.  The Gram matrix is then the exact
    Gram of the stored ``DH`` (csrc/gram_i8.cu).
  * ``distributed=True``: every rank passes ITS shard of the collocation / boundary points and of
    the rows of ``f``; one all-reduce of the packed [Gram | BC Gram | column sums | counts] buffer
    in precompute and one of the projected right-hand sides per solve.
"""
from __future__ import annotations

import os
import pickle
import random
import time

import numpy as np
import torch

from . import _lib, weights as _weights
from .parallel import PackedNormal, allreduce_sum

F, FP = _lib.F, _lib.FP
_JITTER_LADDER = [0.0, 8.0, 128.0, 2048.0, 32768.0, 2.0 ** 20, 2.0 ** 26, 2.0 ** 32]  # x eps x max_diag


def format_input(v, precision=torch.float32, device="cpu", as_array=False, reshape=True):
    """Same contract as the reference helper (utils.py:97-114): lists / ndarrays / tensors are
    moved to ``device`` in ``precision`` and reshaped to columns; python scalars pass through."""
    if isinstance(device, str):
        device = torch.device(device)
    out = []
    for item in v:
        if isinstance(item, (int, float)):
            out.append(item)
            continue
        if isinstance(item, torch.Tensor):
            t = item.to(precision).to(device)
        else:
            t = torch.tensor(np.asarray(item), dtype=precision, device=device)
        out.append(t)
    if reshape:
        out = [t.reshape(-1, 1) if isinstance(t, torch.Tensor) else t for t in out]
    if as_array:
        out = [t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t for t in out]
    return out


def _as_f64_column(item, device):
    """Coordinates go to the kernel in float64 so that a float64 caller loses nothing before the
    Fourier phases are formed (they reach ~40 rad)."""
    if isinstance(item, torch.Tensor):
        return item.detach().to(device=device, dtype=torch.float64).reshape(-1).contiguous()
    return torch.as_tensor(np.asarray(item, dtype=np.float64), device=device).reshape(-1).contiguous()


class Solver:
    def __init__(self, device="cuda", precision=torch.float32, verbose=False, use_weights=True, compile_model=True,
                 lambdas_pde=None, seed=0, *, weights=None, engine=None, chunk_points=0, distributed=False,
                 process_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError(
                "fast_poisson_solver_b200.Solver needs a CUDA device (B200 / sm_100a); there is no CPU fallback. "
                "The reference silently falls back to CPU (solver.py:80); this implementation does not."
            )
        if lambdas_pde is None:
            lambdas_pde = [2 ** -12]
        self.verbose = verbose
        self.precision = precision
        self.device = device
        self.use_weights = use_weights
        self.compile_model = compile_model  # accepted for signature parity, unused
        self.lambdas_pde = list(lambdas_pde)
        self.n_lambdas = len(self.lambdas_pde)
        self.seed = seed
        self.distributed = bool(distributed)
        self.process_group = process_group
        if precision not in (torch.float32, torch.float64):
            raise ValueError("precision must be torch.float32 or torch.float64")

        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)

        self.Ls = np.zeros(self.n_lambdas)
        self.Ls_pde = np.zeros(self.n_lambdas)
        self.Ls_bc = np.zeros(self.n_lambdas)

        self.precompute_path = "precomputed-resources"
        if not os.path.isdir(self.precompute_path):
            os.makedirs(self.precompute_path, exist_ok=True)
        self.precompute_file = os.path.join(self.precompute_path, "default.pkl")

        self.infos = {"model": dict(_weights.MODEL_INFO)}
        self._tdev = torch.device(device)
        if self._tdev.type != "cuda":
            raise RuntimeError(f"device={device!r}: only CUDA devices are supported (no CPU fallback)")
        self._dev_index = self._tdev.index if self._tdev.index is not None else torch.cuda.current_device()
        self._tdev = torch.device("cuda", self._dev_index)

        # ---- weights (solver.py:137-151) --------------------------------------------------------
        if weights is not None:
            sd = _weights.load_state_dict(weights)
        elif use_weights:
            path = os.environ.get("FPS_WEIGHTS", _weights.RESOURCE_WEIGHTS)
            if not os.path.isfile(path):
                raise FileNotFoundError(
                    f"pretrained weights not found at {path}; place the reference's resources/final.pt there, "
                    "set FPS_WEIGHTS, pass weights=<path or state_dict>, or use use_weights=False"
                )
            sd = _weights.load_state_dict(path)
        else:
            sd = _weights.default_init(seed)
        self._host_weights = _weights.HostWeights(sd)

        self._lib = _lib.load()
        if engine is None:
            engine = os.environ.get("FPS_ENGINE", "tcgen05")
        self.engine = {"simt": _lib.ENGINE_SIMT, "tcgen05": _lib.ENGINE_TCGEN05}[engine] if isinstance(engine, str) else int(engine)
        import ctypes as C

        self._ctx = C.c_void_p()
        with torch.cuda.device(self._dev_index):
            _lib.check(self._lib.fps_create(C.byref(self._ctx), self._dev_index, C.byref(self._host_weights.struct),
                                            int(chunk_points), self.engine), "fps_create")
        self._ready = False
        # precision=float64 runs the full-fp64 feature engine (DMMA) and keeps H / DH / outputs in fp64;
        # precision=float32 runs the 3xTF32 tcgen05 engine with fp32 features and fp64 normal equations.
        self._fdtype = torch.float64 if precision == torch.float64 else torch.float32

    def _fn(self, name):
        return getattr(self._lib, name + ("_f64" if self._fdtype == torch.float64 else ""))

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_ctx", None) and self._ctx.value:
                self._lib.fps_destroy(self._ctx)
                self._ctx.value = None
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self._dev_index).cuda_stream

    def set_engine(self, engine):
        self.engine = {"simt": _lib.ENGINE_SIMT, "tcgen05": _lib.ENGINE_TCGEN05}[engine] if isinstance(engine, str) else int(engine)
        _lib.check(self._lib.fps_set_engine(self._ctx, self.engine), "fps_set_engine")

    # ---- kernels ---------------------------------------------------------------------------------
    def _features(self, x64, y64, want_laplace):
        n = x64.numel()
        H = torch.empty((n, FP), dtype=self._fdtype, device=self._tdev)
        DH = torch.empty((n, FP), dtype=self._fdtype, device=self._tdev) if want_laplace else None
        if n:
            _lib.check(self._fn("fps_features")(self._ctx, x64.data_ptr(), y64.data_ptr(), n, H.data_ptr(),
                                              DH.data_ptr() if want_laplace else None, FP, self._stream()),
                       "fps_features")
        return H, DH

    def _allreduce(self, t):
        allreduce_sum(t, self.process_group)

    def _factor(self):
        """Assemble LHS per lambda and Cholesky-factor it (fps_assemble_factor), raising the diagonal
        shift only if the factorisation reports a non-positive pivot."""
        import ctypes as C

        nl = self.n_lambdas
        lam = (C.c_double * nl)(*[float(l) for l in self.lambdas_pde])
        self._L64 = torch.empty((nl, _lib.LROWS, FP), dtype=torch.float64, device=self._tdev)
        info = torch.zeros(nl, dtype=torch.int32, device=self._tdev)
        diag_scale = [float(l) / self.NO * float(self._G64.diagonal()[:F].max()) +
                      float(self._Gbc64.diagonal()[:F].max()) / self.NdO for l in self.lambdas_pde]
        level = [0] * nl
        self.jitter = [0.0] * nl
        for _ in range(len(_JITTER_LADDER)):
            jit = (C.c_double * nl)(*self.jitter)
            _lib.check(self._lib.fps_assemble_factor(self._ctx, self._G64.data_ptr(), self._Gbc64.data_ptr(),
                                                     self._s64.data_ptr(), self.NO, self.NdO, lam, jit, nl,
                                                     self._L64.data_ptr(), info.data_ptr(), self._stream()),
                       "fps_assemble_factor")
            bad = info.cpu().numpy()
            if not bad.any():
                break
            for i in range(nl):
                if bad[i]:
                    level[i] += 1
                    if level[i] >= len(_JITTER_LADDER):
                        raise _lib.FpsError("normal matrix could not be factored even with the largest diagonal shift")
                    self.jitter[i] = _JITTER_LADDER[level[i]] * np.finfo(np.float64).eps * diag_scale[i]
        if self.verbose > 0 and any(self.jitter):
            print("Cholesky needed a diagonal shift:", self.jitter)

    # ---- reference API -----------------------------------------------------------------------------
    def precompute(self, x_pde, y_pde, x_bc, y_bc, name=None, save=False, load=False):
        t0 = time.perf_counter()
        if name is not None:
            self.precompute_file = os.path.join(self.precompute_path, name + ".pkl")
        if self.distributed and ".rank" not in os.path.basename(self.precompute_file):
            # row-sharded state: one file per rank
            import torch.distributed as dist

            self.precompute_file = self.precompute_file[:-4] + f".rank{dist.get_rank(self.process_group)}.pkl"
        with torch.cuda.device(self._dev_index):
            x64 = [_as_f64_column(v, self._tdev) for v in (x_pde, y_pde, x_bc, y_bc)]
            self._xy64 = x64
            self.x_pde, self.y_pde, self.x_bc, self.y_bc = [t.to(self.precision).reshape(-1, 1) for t in x64]
            self.x_tot = torch.cat([self.x_pde, self.x_bc], dim=0)
            self.y_tot = torch.cat([self.y_pde, self.y_bc], dim=0)
            if self.x_tot.numel():
                assert torch.min(self.x_tot) >= 0 and torch.max(self.x_tot) <= 1, \
                    "x coordinates should be in [0, 1]. Please rescale."
                assert torch.min(self.y_tot) >= 0 and torch.max(self.y_tot) <= 1, \
                    "y coordinates should be in [0, 1]. Please rescale."

            # drop the previous domain's feature matrices first: at 16M points they are 47 GB each
            for name in ("H", "DH", "Dht", "H_bc", "Ht_bc", "_Hp", "_DHp", "_Hbcp", "u_pred", "u_pde_pred", "u_bc_pred",
                         "f_pred", "f", "bc"):
                if hasattr(self, name):
                    setattr(self, name, None)
            use_file = bool(load) and os.path.isfile(self.precompute_file)
            if self.distributed and load:
                # every rank must take the same branch (the recompute branch contains collectives)
                import torch.distributed as dist

                flag = torch.tensor([1.0 if use_file else 0.0], device=self._tdev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.process_group)
                use_file = flag.item() > 0.5
            if use_file:
                self._load_precomputed()
            else:
                if self._fdtype == torch.float32 and x64[0].numel():
                    # fit the fp16 operand scales of the tensor-core engine to this network / domain (1 ms)
                    # on an evenly strided sample of the collocation points (grids come sorted)
                    n_all = x64[0].numel()
                    if n_all > 512:
                        pick = torch.linspace(0, n_all - 1, 512, device=self._tdev).long()
                        xs, ys = x64[0][pick].contiguous(), x64[1][pick].contiguous()
                    else:
                        xs, ys = x64[0], x64[1]
                    _lib.check(self._lib.fps_calibrate_scales(self._ctx, xs.data_ptr(), ys.data_ptr(), xs.numel(),
                                                              self._stream()), "fps_calibrate_scales")
                self._Hp, self._DHp = self._features(x64[0], x64[1], True)      # solver.py:162-166
                self._Hbcp, _ = self._features(x64[2], x64[3], False)           # solver.py:173-175
                packed = PackedNormal(self._tdev)
                pack = packed.buf
                self._G64, self._Gbc64, self._s64 = packed.G, packed.Gbc, packed.s
                st = self._stream()
                n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
                if n_loc:
                    _lib.check(self._fn("fps_gram_accum")(self._ctx, self._DHp.data_ptr(), n_loc, FP,
                                                        self._G64.data_ptr(), st), "fps_gram_accum")  # solver.py:167
                if nbc_loc:
                    _lib.check(self._fn("fps_gram_accum")(self._ctx, self._Hbcp.data_ptr(), nbc_loc, FP,
                                                        self._Gbc64.data_ptr(), st), "fps_gram_accum(bc)")
                    _lib.check(self._fn("fps_colsum_accum")(self._ctx, self._Hbcp.data_ptr(), nbc_loc, FP,
                                                          self._s64.data_ptr(), st), "fps_colsum_accum")  # solver.py:176-178
                packed.set_counts(n_loc, nbc_loc)
                if self.distributed:
                    packed.allreduce(self.process_group)  # the path's one exchange step
                    self.NO, self.NdO = packed.counts()
                else:
                    self.NO, self.NdO = n_loc, nbc_loc
                self._pack = pack
                self._factor()                                                  # solver.py:187-199 + LU of :305
                if save:
                    self._save_precomputed()
            self._publish()
            torch.cuda.synchronize(self._dev_index)
        self._ready = True
        t3 = time.perf_counter()
        if self.verbose > 0:
            print("\nPre-Computing Successful:", t3 - t0, "seconds")

    def _publish(self):
        """Expose reference attribute names as views / small casts of the device state."""
        self.H = self._Hp[:, :F]
        self.DH = self._DHp[:, :F]
        self.Dht = self.DH.t()
        self.H_bc = self._Hbcp[:, :F]
        self.Ht_bc = self.H_bc.t()
        self.DHtDH = self._G64[:F, :F].to(self.precision)
        self.Ht_bc_ones = self._s64[:F].to(self.precision).reshape(1, F)

    @property
    def LHSs(self):
        """(n_lambda, 700, 700) normal matrices as precompute_LHS_RHS builds them (solver.py:195-196)."""
        G, Gbc, s = self._G64[:F, :F], self._Gbc64[:F, :F], self._s64[:F]
        lam = torch.tensor(self.lambdas_pde, dtype=torch.float64, device=self._tdev).view(-1, 1, 1)
        bc = (Gbc - torch.outer(s, s) / self.NdO) / self.NdO
        return (lam / self.NO * G[None] + bc[None]).to(self.precision)

    @property
    def RHSs(self):
        """(n_lambda, 700, N) = lambda/N * DH^T (solver.py:197); materialised only on request."""
        lam = torch.tensor(self.lambdas_pde, dtype=self.precision, device=self._tdev).view(-1, 1, 1)
        return lam / self.NO * self.Dht[None]

    # ---- precomputed-state persistence (solver.py:180-185, :201-207) ------------------------------
    def _save_precomputed(self):
        payload = {
            "format": "fps_b200/2",
            "H": self._Hp.cpu(), "DH": self._DHp.cpu(), "H_bc": self._Hbcp.cpu(), "pack": self._pack.cpu(),
            "L64": self._L64.cpu(), "NO": self.NO, "NdO": self.NdO, "lambdas": self.lambdas_pde, "jitter": self.jitter,
        }
        with open(self.precompute_file, "wb") as fh:
            pickle.dump(payload, fh, protocol=pickle.HIGHEST_PROTOCOL)
        if self.verbose > 0:
            print("Pre-Computed stored.")

    def _load_precomputed(self):
        with open(self.precompute_file, "rb") as fh:
            p = pickle.load(fh)
        if not isinstance(p, dict) or p.get("format") != "fps_b200/2":
            raise _lib.FpsError(f"{self.precompute_file} was not written by this implementation")
        self._Hp, self._DHp, self._Hbcp = (p[k].to(self._tdev) for k in ("H", "DH", "H_bc"))
        pack = p["pack"].to(self._tdev)
        self._pack = pack
        self._G64 = pack[: FP * FP].view(FP, FP)
        self._Gbc64 = pack[FP * FP: 2 * FP * FP].view(FP, FP)
        self._s64 = pack[2 * FP * FP: 2 * FP * FP + FP]
        self.NO, self.NdO = p["NO"], p["NdO"]
        if list(p["lambdas"]) == list(self.lambdas_pde):
            self._L64, self.jitter = p["L64"].to(self._tdev), p["jitter"]
        else:
            self._factor()
        if self.verbose > 0:
            print("Pre-Computed data_utils loaded from storage.")

    # ---- solve -------------------------------------------------------------------------------------
    def _canon_rhs(self, f, bc):
        """Single RHS: anything with N (resp. N_bc) elements, flattened like utils.py:110-111.
        Batched extension: f (N, B), bc (N_bc, B) or B constants."""
        n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
        ft = f if isinstance(f, torch.Tensor) else torch.as_tensor(np.asarray(f))
        bt = bc if isinstance(bc, torch.Tensor) else torch.as_tensor(np.asarray(bc))
        ft = ft.to(self._tdev)
        bt = bt.to(self._tdev)
        if ft.numel() == n_loc:
            B = 1
            ft = ft.reshape(n_loc, 1)
        elif ft.dim() == 2 and ft.shape[0] == n_loc:
            B = ft.shape[1]
        else:
            raise ValueError(f"f has shape {tuple(ft.shape)}; expected {n_loc} values or ({n_loc}, B)")
        if B == 1:
            bt = bt.reshape(-1, 1)
        elif bt.dim() == 2 and bt.shape == (nbc_loc, B):
            pass
        elif bt.numel() == B:
            bt = bt.reshape(1, B).expand(nbc_loc, B)
        else:
            raise ValueError(f"bc has shape {tuple(bt.shape)}; expected ({nbc_loc}, {B}) or {B} constants")
        return ft.to(self._fdtype).contiguous(), bt, B

    def solve(self, f, bc):
        if not self._ready:
            raise RuntimeError("call precompute() before solve()")
        with torch.cuda.device(self._dev_index):
            F32, BC, B = self._canon_rhs(f, bc)
            self.f, self.bc = F32.to(self.precision), BC.to(self.precision)
            st = self._stream()
            n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
            torch.cuda.synchronize(self._dev_index)
            t0 = time.perf_counter()

            R = torch.zeros((FP, B), dtype=torch.float64, device=self._tdev)
            if n_loc:
                _lib.check(self._fn("fps_project_rhs")(self._ctx, self._DHp.data_ptr(), n_loc, FP, F32.data_ptr(),
                                                     F32.stride(0), B, R.data_ptr(), B, st), "fps_project_rhs")  # :301
            sum_bc = BC.to(torch.float64).sum(dim=0).contiguous()                                               # :302
            if self.distributed:
                self._allreduce(R)
                self._allreduce(sum_bc)
            U = torch.empty((n_loc + nbc_loc, B), dtype=self._fdtype, device=self._tdev)
            Fp = torch.empty((n_loc, B), dtype=self._fdtype, device=self._tdev)
            W = torch.empty((FP, B), dtype=torch.float64, device=self._tdev)
            bias = torch.empty(B, dtype=torch.float64, device=self._tdev)
            best = None
            for i, lam in enumerate(self.lambdas_pde):
                _lib.check(self._lib.fps_solve_factored(self._ctx, self._L64[i].data_ptr(), R.data_ptr(), B, B,
                                                        float(lam) / self.NO, self._s64.data_ptr(), sum_bc.data_ptr(),
                                                        self.NdO, W.data_ptr(), bias.data_ptr(), st),
                           "fps_solve_factored")                                                                # :305-306
                self._reconstruct(W, bias, B, U, Fp, st)                                                         # :327-331
                if self.n_lambdas > 1:
                    # loss sweep of solver.py:309-317 (per right-hand side)
                    l_pde = ((Fp.double() - F32.double()) ** 2).sum(dim=0)
                    l_bc = ((U[n_loc:].double() - BC.double()) ** 2).sum(dim=0)
                    cnt = torch.tensor([float(n_loc), float(nbc_loc)], dtype=torch.float64, device=self._tdev)
                    if self.distributed:
                        self._allreduce(l_pde)
                        self._allreduce(l_bc)
                        self._allreduce(cnt)
                    l_pde, l_bc = l_pde / cnt[0], l_bc / cnt[1]
                    self.Ls_pde[i], self.Ls_bc[i] = l_pde[0].item(), l_bc[0].item()
                    self.Ls[i] = self.Ls_pde[i] + self.Ls_bc[i]
                    tot = l_pde + l_bc
                    if best is None:
                        best = (tot.clone(), U.clone(), Fp.clone(), W.clone(), bias.clone(),
                                torch.zeros(B, dtype=torch.long, device=self._tdev))
                    else:
                        take = tot < best[0]
                        best[0][take] = tot[take]
                        best[1][:, take] = U[:, take]
                        best[2][:, take] = Fp[:, take]
                        best[3][:, take] = W[:, take]
                        best[4][take] = bias[take]
                        best[5][take] = i
            if self.n_lambdas > 1:
                _, U, Fp, W, bias, which = best
                self.lambda_pde = self.lambdas_pde[int(which[0].item())]
                self.lambda_index = which
            else:
                self.lambda_pde = self.lambdas_pde[0]

            self._W64, self._bias64 = W, bias
            self.w_out = W[:F].to(self.precision)
            self.bias = bias.to(self.precision).reshape(1, B)
            self.u_pred = U.to(self.precision)
            self.f_pred = Fp.to(self.precision)
            self.u_pde_pred = self.u_pred[:n_loc]
            self.u_bc_pred = self.u_pred[n_loc:]
            torch.cuda.synchronize(self._dev_index)
            t1 = time.perf_counter()
        if self.verbose > 0:
            print("\nRun Successful:", t1 - t0, "seconds\n")
        return self.u_pred, self.u_pde_pred, self.u_bc_pred, self.f_pred, t1 - t0

    def evaluate(self, x, y):
        """Extension: the last solution u(x, y) = H(x, y) w + b at arbitrary points of [0,1]^2, (n, B)."""
        if not hasattr(self, "_W64"):
            raise RuntimeError("call solve() before evaluate()")
        with torch.cuda.device(self._dev_index):
            x64, y64 = _as_f64_column(x, self._tdev), _as_f64_column(y, self._tdev)
            Hq, _ = self._features(x64, y64, False)
            B = self._W64.shape[1]
            out = torch.empty((x64.numel(), B), dtype=self._fdtype, device=self._tdev)
            _lib.check(self._fn("fps_reconstruct")(self._ctx, Hq.data_ptr(), x64.numel(), FP, self._W64.data_ptr(), B,
                                                   self._bias64.data_ptr(), B, out.data_ptr(), B, self._stream()),
                       "fps_reconstruct(evaluate)")
            return out.to(self.precision)

    def _reconstruct(self, W, bias, B, U, Fp, st):
        n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
        lib, ctx = self._lib, self._ctx
        if n_loc:
            _lib.check(self._fn("fps_reconstruct")(ctx, self._Hp.data_ptr(), n_loc, FP, W.data_ptr(), B, bias.data_ptr(), B,
                                           U.data_ptr(), B, st), "fps_reconstruct(u_pde)")
            _lib.check(self._fn("fps_reconstruct")(ctx, self._DHp.data_ptr(), n_loc, FP, W.data_ptr(), B, None, B,
                                           Fp.data_ptr(), B, st), "fps_reconstruct(f)")
        if nbc_loc:
            _lib.check(self._fn("fps_reconstruct")(ctx, self._Hbcp.data_ptr(), nbc_loc, FP, W.data_ptr(), B, bias.data_ptr(), B,
                                           U[n_loc:].data_ptr(), B, st), "fps_reconstruct(u_bc)")
