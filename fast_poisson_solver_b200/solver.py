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
  * For N >= 16384 (float32 mode) ``precompute`` stores ``H`` and ``DH`` on fixed-point grids: ``H`` on multiples of
    2^-30 (|H| <= 1), ``DH`` on a 31-bit grid per feature column (elements far below the column maximum move by
    2^-31..2^-29 of it, under 1 % of one fp32 ulp of the maximum; all others are unchanged).  The Gram matrix is the exact
    Gram of the stored ``DH`` (csrc/gram_i8.cu).  ``solve`` never modifies precomputed state.
  * ``distributed=True``: every rank passes ITS shard of the collocation / boundary points and of
    the rows of ``f``; one all-reduce of the packed [Gram | BC Gram | column sums | counts] buffer
    in precompute and one of the projected right-hand sides per solve.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import random
import time

import numpy as np
import torch

from . import _lib, weights as _weights
from .parallel import Communicator, PackedNormal

F, FP = _lib.F, _lib.FP
_JITTER_LADDER = [0.0, 8.0, 128.0, 2048.0, 32768.0, 2.0 ** 20, 2.0 ** 26, 2.0 ** 32]  # x eps x max_diag
_FORMAT = "fps_b200/3"
_MIN_B_EXACT = 32   # right-hand sides from which projection / reconstruction use the int8 engine


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
                 process_group=None, cache_planes=None):
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
        self.weights_sha256 = self._host_weights.sha256()

        self._lib = _lib.load()
        if engine is None:
            engine = os.environ.get("FPS_ENGINE", "tcgen05")
        self.engine = {"simt": _lib.ENGINE_SIMT, "tcgen05": _lib.ENGINE_TCGEN05}[engine] if isinstance(engine, str) else int(engine)

        self._ctx = C.c_void_p()
        with torch.cuda.device(self._dev_index):
            _lib.check(self._lib.fps_create(C.byref(self._ctx), self._dev_index, C.byref(self._host_weights.struct),
                                            int(chunk_points), self.engine), "fps_create")
        self._ready = False
        # precision=float64 runs the full-fp64 feature engine (DMMA) and keeps H / DH / outputs in fp64;
        # precision=float32 runs the tcgen05 engine with fp32-grade features and fp64-grade normal equations.
        self._fdtype = torch.float64 if precision == torch.float64 else torch.float32
        # digit planes of the exact int8 engine are kept between calls when memory allows (None = decide per problem:
        # at most a quarter of the free device memory; FPS_CACHE_PLANES=0/1 or cache_planes=False/True force it)
        env = os.environ.get("FPS_CACHE_PLANES")
        self._cache_planes = cache_planes if cache_planes is not None else (None if env is None else env != "0")
        self._cache_ok_bytes = None
        self._fixed_cal = None      # a calibration state imposed by set_calibration() (None: measured per precompute)
        self._exact_min = int(self._lib.fps_exact_min_rows())
        self._reserved = (0, 0)
        self._comm = None

    def _fn(self, name):
        return getattr(self._lib, name + ("_f64" if self._fdtype == torch.float64 else ""))

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_ctx", None) and self._ctx.value:
                self._lib.fps_destroy(self._ctx)   # restores the caller's current device itself
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

    def _exact(self, n):
        """True when n rows go through the cached exact int8 engine (float32 mode, n >= fps_exact_min_rows())."""
        return self._fdtype == torch.float32 and n >= self._exact_min

    def _want_cache(self, nbytes):
        if self._cache_planes is not None:
            return bool(self._cache_planes)
        # cudaMemGetInfo costs ~0.3 ms (measured: the largest idle gap of a step), so a positive answer for a size is
        # remembered; if memory got scarce since, the allocation itself fails and the caller falls back
        if self._cache_ok_bytes is not None and nbytes <= self._cache_ok_bytes:
            return True
        free, _ = torch.cuda.mem_get_info(self._dev_index)
        free += torch.cuda.memory_reserved(self._dev_index) - torch.cuda.memory_allocated(self._dev_index)
        ok = nbytes <= free // 4
        if ok:
            self._cache_ok_bytes = nbytes
        return ok

    def _alloc_planes(self, nbytes):
        try:
            return torch.empty(nbytes, dtype=torch.int8, device=self._tdev)
        except torch.cuda.OutOfMemoryError:
            return None

    def _reserve(self, n, B):
        """Size the library's own scratch for (n rows, B right-hand sides) so that no call allocates mid-stream."""
        if self._exact(n) and (n > self._reserved[0] or B > self._reserved[1]):
            self._reserved = (max(n, self._reserved[0]), max(B, self._reserved[1]))
            _lib.check(self._lib.fps_reserve_workspace(self._ctx, self._reserved[0], self._reserved[1]),
                       "fps_reserve_workspace")

    def _allreduce(self, t):
        """The path's exchange step: an in-place fp64 sum over the ranks (csrc/comm.cu through parallel.Communicator)."""
        if self._comm is None:
            self._comm = Communicator(self._lib, self._dev_index, self.process_group)
        self._comm.allreduce(t, self._stream())

    def _enqueue_factor(self, jitter):
        nl = self.n_lambdas
        lam = (C.c_double * nl)(*[float(l) for l in self.lambdas_pde])
        jit = (C.c_double * nl)(*jitter)
        _lib.check(self._lib.fps_assemble_factor_packed(self._ctx, self._pack.data_ptr(), lam, jit, nl,
                                                        self._L64.data_ptr(), self._info.data_ptr(), self._stream()),
                   "fps_assemble_factor_packed")

    def _factor(self):
        """Assemble LHS per lambda and Cholesky-factor it (solver.py:187-199 + the LU of :305).  The first attempt is
        enqueued without any host round trip; `_finish_factor` looks at the pivot report once the stream is drained."""
        nl = self.n_lambdas
        self._L64 = torch.empty((nl, _lib.LROWS, FP), dtype=torch.float64, device=self._tdev)
        self._info = torch.zeros(nl, dtype=torch.int32, device=self._tdev)
        self.jitter = [0.0] * nl
        self._enqueue_factor(self.jitter)

    def _finish_factor(self):
        """Raise the diagonal shift only if the factorisation reported a non-positive pivot (rank-deficient systems,
        e.g. fewer equations than features).  Slow path: host-driven ladder."""
        nl = self.n_lambdas
        bad = self._info.cpu().numpy()
        if not bad.any():
            return
        diag_scale = [float(l) / self.NO * float(self._G64.diagonal()[:F].max()) +
                      float(self._Gbc64.diagonal()[:F].max()) / self.NdO for l in self.lambdas_pde]
        level = [0] * nl
        while bad.any():
            for i in range(nl):
                if bad[i]:
                    level[i] += 1
                    if level[i] >= len(_JITTER_LADDER):
                        raise _lib.FpsError("normal matrix could not be factored even with the largest diagonal shift")
                    self.jitter[i] = _JITTER_LADDER[level[i]] * np.finfo(np.float64).eps * diag_scale[i]
            self._enqueue_factor(self.jitter)
            bad = self._info.cpu().numpy()
        if self.verbose > 0:
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
            # range check of solver.py:244-247: evaluated on the device now, read back (and raised) at the end of
            # precompute, where the stream is drained anyway - no host round trip in the middle of the pipeline
            rng = None
            if self.x_tot.numel():
                rng = torch.stack(torch.aminmax(self.x_tot) + torch.aminmax(self.y_tot)).double()

            # drop the previous domain's feature matrices first: at 16M points they are 47 GB each
            for attr in ("H", "DH", "Dht", "H_bc", "Ht_bc", "_Hp", "_DHp", "_Hbcp", "u_pred", "u_pde_pred", "u_bc_pred",
                         "f_pred", "f", "bc", "_gplanes", "_rplanes_h", "_rplanes_dh", "_W64", "_bias64"):
                if hasattr(self, attr):
                    setattr(self, attr, None)
            use_file = bool(load) and os.path.isfile(self.precompute_file)
            if self.distributed and load:
                # every rank must take the same branch (the recompute branch contains collectives)
                import torch.distributed as dist

                flag = torch.tensor([1.0 if use_file else 0.0], device=self._tdev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.process_group)
                use_file = flag.item() > 0.5
            loaded = False
            if use_file:
                loaded = self._load_precomputed(x64[0].numel(), x64[2].numel())
                if self.distributed:
                    import torch.distributed as dist

                    flag = torch.tensor([1.0 if loaded else 0.0], device=self._tdev)
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.process_group)
                    loaded = flag.item() > 0.5
            if not loaded:
                self._compute(x64)
                if save:
                    self._save_precomputed()
            self._publish()
            torch.cuda.synchronize(self._dev_index)
            if rng is not None:
                lo_x, hi_x, lo_y, hi_y = rng.tolist()
                assert lo_x >= 0 and hi_x <= 1, "x coordinates should be in [0, 1]. Please rescale."
                assert lo_y >= 0 and hi_y <= 1, "y coordinates should be in [0, 1]. Please rescale."
            if not loaded or self._refactor_pending:
                self._finish_factor()
        self._ready = True
        t3 = time.perf_counter()
        if self.verbose > 0:
            print("\nPre-Computing Successful:", t3 - t0, "seconds")

    def _calibrate(self, x64):
        """Fit the fp16 operand scales of the tensor-core engine (and the fixed-point grid of DH) to this network /
        domain on an evenly strided sample of 512 collocation points (grids come sorted).  With `distributed=True` the
        states of all ranks are combined (one small all-reduce): the calibration belongs to the problem, not to the
        shard, so a point's features do not depend on the rank it was assigned to."""
        if self._fdtype != torch.float32:
            return
        if self._fixed_cal is not None:
            self._apply_calibration(self._fixed_cal)
            return
        n_all = x64[0].numel()
        if n_all:
            if n_all > 512:
                pick = torch.linspace(0, n_all - 1, 512, device=self._tdev).long()
                xs, ys = x64[0][pick].contiguous(), x64[1][pick].contiguous()
            else:
                xs, ys = x64[0], x64[1]
            _lib.check(self._lib.fps_calibrate_scales(self._ctx, xs.data_ptr(), ys.data_ptr(), xs.numel(), self._stream()),
                       "fps_calibrate_scales")
        if self.distributed:
            import torch.distributed as dist

            # MIN of the power-of-two scales and of the gain switch, MAX of the DH maxima, as ONE max-reduction of
            # [-scales | -gain_on | dh maxima]; a rank without points contributes the neutral element
            st = self.calibration_state() if n_all else {"scales": [3.0e38] * 24, "gain_on": -1, "dh_cmax": np.zeros(_lib.FPW)}
            keep_gain = 0.0 if st["gain_on"] == 0 else 1.0          # switched off as soon as one rank found it harmful
            v = np.concatenate([-np.asarray(st["scales"], dtype=np.float64), [-keep_gain],
                                np.asarray(st["dh_cmax"], dtype=np.float64)])
            t = torch.as_tensor(v, device=self._tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.process_group)
            v = t.cpu().numpy()
            if -v[0] < 1.0e38:      # at least one rank had points
                gain = 0 if -v[24] < 0.5 else (1 if st["gain_on"] >= 0 else -1)
                self._apply_calibration({"scales": (-v[:24]).tolist(), "gain_on": gain, "dh_cmax": v[25:]})

    def calibration_state(self):
        """The calibration the last precompute ran with: 6 x 4 power-of-two operand scales (layer input x stream), the
        accumulator-gain switch (1 / 0; -1 = not checked) and max |DH| per feature (768 floats, zero = none)."""
        sc, cal, dh = (C.c_float * 24)(), (C.c_double * 5)(), (C.c_float * _lib.FPW)()
        _lib.check(self._lib.fps_get_scales(self._ctx, sc), "fps_get_scales")
        _lib.check(self._lib.fps_get_calibration(self._ctx, cal), "fps_get_calibration")
        _lib.check(self._lib.fps_get_dh_calibration(self._ctx, dh), "fps_get_dh_calibration")
        return {"scales": [float(v) for v in sc], "gain_on": int(cal[4]), "dh_cmax": np.array(dh, dtype=np.float32)}

    def set_calibration(self, state):
        """Impose a calibration state (from `calibration_state()` of another Solver) on every later precompute instead of
        measuring one; None returns to measuring.  Two Solvers with the same weights and the same calibration produce
        bit-identical features for the same point (e.g. a single-GPU run checked against a sharded one)."""
        self._fixed_cal = state

    def _apply_calibration(self, state):
        sc = (C.c_float * 24)(*[float(v) for v in state["scales"]])
        _lib.check(self._lib.fps_set_calibration(self._ctx, sc, int(state.get("gain_on", -1))), "fps_set_calibration")
        dh = (C.c_float * _lib.FPW)(*[float(v) for v in np.asarray(state["dh_cmax"], dtype=np.float32)])
        _lib.check(self._lib.fps_set_dh_calibration(self._ctx, dh), "fps_set_dh_calibration")

    def _compute(self, x64):
        self._refactor_pending = False
        st = self._stream()
        lib, ctx = self._lib, self._ctx
        n_loc, nbc_loc = x64[0].numel(), x64[2].numel()
        self._calibrate(x64)
        packed = PackedNormal(self._tdev)
        self._pack = packed.buf
        self._G64, self._Gbc64, self._s64 = packed.G, packed.Gbc, packed.s
        self._xop_h = self._xop_dh = self._gplanes = self._rplanes_h = self._rplanes_dh = None
        if self._exact(n_loc):
            # cached exact-engine path: H / DH leave the feature kernel on their fixed-point grids, together with the
            # digit planes of DH that the Gram and every later batched projection read (solver.py:162-167)
            self._reserve(n_loc, 0)
            self._Hp = torch.empty((n_loc, FP), dtype=torch.float32, device=self._tdev)
            self._DHp = torch.empty((n_loc, FP), dtype=torch.float32, device=self._tdev)
            xops = torch.zeros((2, _lib.XOP_BYTES // 4), dtype=torch.int32, device=self._tdev)
            self._xop_h, self._xop_dh = xops[0], xops[1]
            gbytes = int(lib.fps_gram_planes_bytes(n_loc))
            self._gplanes = self._alloc_planes(gbytes) if self._want_cache(gbytes) else None
            if self._gplanes is not None:
                written = C.c_int(0)
                _lib.check(lib.fps_features_x(ctx, x64[0].data_ptr(), x64[1].data_ptr(), n_loc, self._Hp.data_ptr(),
                                              self._DHp.data_ptr(), FP, self._xop_h.data_ptr(), self._xop_dh.data_ptr(),
                                              self._gplanes.data_ptr(), C.byref(written), st), "fps_features_x")
                _lib.check(lib.fps_gram_accum_x(ctx, self._DHp.data_ptr(), n_loc, FP, self._xop_dh.data_ptr(),
                                                self._gplanes.data_ptr(), written.value, self._G64.data_ptr(), st),
                           "fps_gram_accum_x")
            else:
                # not enough memory to keep the planes: same arithmetic through the library's own pass workspace
                _lib.check(lib.fps_features_x(ctx, x64[0].data_ptr(), x64[1].data_ptr(), n_loc, self._Hp.data_ptr(),
                                              self._DHp.data_ptr(), FP, self._xop_h.data_ptr(), self._xop_dh.data_ptr(),
                                              None, None, st), "fps_features_x")
                _lib.check(lib.fps_gram_accum(ctx, self._DHp.data_ptr(), n_loc, FP, self._G64.data_ptr(), st),
                           "fps_gram_accum")
                self._xop_h = self._xop_dh = None    # later calls take the uncached entry points (same grids, same results)
        else:
            self._Hp, self._DHp = self._features(x64[0], x64[1], True)      # solver.py:162-166
            if n_loc:
                _lib.check(self._fn("fps_gram_accum")(ctx, self._DHp.data_ptr(), n_loc, FP, self._G64.data_ptr(), st),
                           "fps_gram_accum")                                # solver.py:167
        self._Hbcp, _ = self._features(x64[2], x64[3], False)               # solver.py:173-175
        if nbc_loc:
            _lib.check(self._fn("fps_gram_accum")(ctx, self._Hbcp.data_ptr(), nbc_loc, FP, self._Gbc64.data_ptr(), st),
                       "fps_gram_accum(bc)")
            _lib.check(self._fn("fps_colsum_accum")(ctx, self._Hbcp.data_ptr(), nbc_loc, FP, self._s64.data_ptr(), st),
                       "fps_colsum_accum")                                  # solver.py:176-178
        packed.set_counts(n_loc, nbc_loc)
        if self.distributed:
            self._allreduce(packed.buf)            # the path's one exchange step
            self._counts_dev = packed.buf[-2:]
            self.NO = self.NdO = None              # read back with the final synchronise (see _publish)
        else:
            self.NO, self.NdO = n_loc, nbc_loc
        self._factor()                                                      # solver.py:187-199 + LU of :305

    def _publish(self):
        """Expose reference attribute names as views / small casts of the device state."""
        if self.NO is None:
            c = self._counts_dev.cpu()
            self.NO, self.NdO = int(round(c[0].item())), int(round(c[1].item()))
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
        torch.cuda.synchronize(self._dev_index)
        self._finish_factor()
        if self.NO is None:
            c = self._counts_dev.cpu()
            self.NO, self.NdO = int(round(c[0].item())), int(round(c[1].item()))
        scales = (C.c_float * 24)()
        _lib.check(self._lib.fps_get_scales(self._ctx, scales), "fps_get_scales")
        payload = {
            "format": _FORMAT, "dtype": str(self._fdtype), "FP": FP, "n": self._Hp.shape[0], "n_bc": self._Hbcp.shape[0],
            "weights_sha256": self.weights_sha256, "act_scales": list(scales), "gain_on": self._gain_on(),
            "H": self._Hp.cpu(), "DH": self._DHp.cpu(), "H_bc": self._Hbcp.cpu(), "pack": self._pack.cpu(),
            "xop_h": None if self._xop_h is None else self._xop_h.cpu(),
            "xop_dh": None if self._xop_dh is None else self._xop_dh.cpu(),
            "L64": self._L64.cpu(), "NO": self.NO, "NdO": self.NdO, "lambdas": self.lambdas_pde, "jitter": self.jitter,
        }
        with open(self.precompute_file, "wb") as fh:
            pickle.dump(payload, fh, protocol=pickle.HIGHEST_PROTOCOL)
        if self.verbose > 0:
            print("Pre-Computed stored.")

    def _gain_on(self):
        cal = (C.c_double * 5)()
        _lib.check(self._lib.fps_get_calibration(self._ctx, cal), "fps_get_calibration")
        return int(cal[4])

    def _load_precomputed(self, n_loc, nbc_loc):
        """Returns False (-> recompute, as the reference does for a missing file) when the file does not describe this
        Solver's problem: foreign format, other precision / weights / point counts."""
        try:
            with open(self.precompute_file, "rb") as fh:
                p = pickle.load(fh)
        except Exception:
            return False
        ok = (isinstance(p, dict) and p.get("format") == _FORMAT and p.get("dtype") == str(self._fdtype)
              and p.get("FP") == FP and p.get("n") == n_loc and p.get("n_bc") == nbc_loc
              and p.get("weights_sha256") == self.weights_sha256)
        if ok:
            ok = (tuple(p["H"].shape) == (n_loc, FP) and tuple(p["DH"].shape) == (n_loc, FP)
                  and tuple(p["H_bc"].shape) == (nbc_loc, FP) and p["H"].dtype == self._fdtype
                  and p["pack"].numel() == PackedNormal.SIZE)
        if not ok:
            if self.verbose > 0:
                print(f"{self.precompute_file} does not match this Solver (format / precision / weights / points): recomputing.")
            return False
        self._Hp, self._DHp, self._Hbcp = (p[k].to(self._tdev) for k in ("H", "DH", "H_bc"))
        pack = p["pack"].to(self._tdev)
        self._pack = pack
        self._G64 = pack[: FP * FP].view(FP, FP)
        self._Gbc64 = pack[FP * FP: 2 * FP * FP].view(FP, FP)
        self._s64 = pack[2 * FP * FP: 2 * FP * FP + FP]
        self.NO, self.NdO = p["NO"], p["NdO"]
        self._xop_h = None if p.get("xop_h") is None else p["xop_h"].to(self._tdev)
        self._xop_dh = None if p.get("xop_dh") is None else p["xop_dh"].to(self._tdev)
        self._gplanes = self._rplanes_h = self._rplanes_dh = None   # rebuilt on demand from the stored (snapped) matrices
        # the operand scales the stored H was produced with (evaluate() must reproduce the same features)
        sc = (C.c_float * 24)(*p["act_scales"])
        _lib.check(self._lib.fps_set_calibration(self._ctx, sc, int(p.get("gain_on", -1))), "fps_set_calibration")
        self._refactor_pending = False
        if list(p["lambdas"]) == list(self.lambdas_pde) and tuple(p["L64"].shape[1:]) == (_lib.LROWS, FP):
            self._L64, self.jitter = p["L64"].to(self._tdev), p["jitter"]
        else:
            self._factor()
            self._refactor_pending = True
        if self.verbose > 0:
            print("Pre-Computed data_utils loaded from storage.")
        return True

    # ---- solve -------------------------------------------------------------------------------------
    def _canon_rhs(self, f, bc):
        """Single RHS: anything with N (resp. N_bc) elements, flattened like utils.py:110-111.
        Batched extension: f (N, B), bc (N_bc, B) or B constants (one constant Dirichlet value per right-hand side)."""
        n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
        ft = f if isinstance(f, torch.Tensor) else torch.as_tensor(np.asarray(f))
        bt = bc if isinstance(bc, torch.Tensor) else torch.as_tensor(np.asarray(bc))
        ft = ft.to(self._tdev)
        bt = bt.to(self._tdev)
        if ft.numel() == n_loc and not (ft.dim() == 2 and ft.shape[0] == n_loc and ft.shape[1] != 1):
            B = 1
            ft = ft.reshape(n_loc, 1)
        elif ft.dim() == 2 and ft.shape[0] == n_loc:
            B = ft.shape[1]
        else:
            raise ValueError(f"f has shape {tuple(ft.shape)}; expected {n_loc} values or ({n_loc}, B)")
        if bt.numel() == nbc_loc * B and (B == 1 or (bt.dim() == 2 and bt.shape == (nbc_loc, B))):
            bt = bt.reshape(nbc_loc, B)
        elif bt.numel() == B:
            bt = bt.reshape(1, B).expand(nbc_loc, B)        # B constants (B = 1: a scalar boundary value)
        else:
            raise ValueError(f"bc has shape {tuple(bt.shape)}; expected {nbc_loc} values"
                             + (f", ({nbc_loc}, {B})" if B > 1 else "") + f" or {B} constant(s)")
        return ft.to(self._fdtype).contiguous(), bt, B

    def _ensure_recon_planes(self):
        """Feature-major digit planes of H and DH for the batched reconstruction: built once per precompute, on the
        first batched solve (a pure cache: H and DH already lie on their grids, nothing is modified)."""
        if self._rplanes_h is not None or self._xop_h is None:
            return
        n_loc = self._DHp.shape[0]
        nbytes = int(self._lib.fps_recon_planes_bytes(n_loc))
        if not self._want_cache(2 * nbytes):
            self._rplanes_h = self._rplanes_dh = False
            return
        ph, pdh = self._alloc_planes(nbytes), self._alloc_planes(nbytes)
        if ph is None or pdh is None:
            self._rplanes_h = self._rplanes_dh = False
            return
        st = self._stream()
        for A, xop, pl in ((self._Hp, self._xop_h, ph), (self._DHp, self._xop_dh, pdh)):
            _lib.check(self._lib.fps_recon_planes_build(self._ctx, A.data_ptr(), n_loc, FP, xop.data_ptr(), 0,
                                                        pl.data_ptr(), st), "fps_recon_planes_build")
        self._rplanes_h, self._rplanes_dh = ph, pdh

    def solve(self, f, bc):
        if not self._ready:
            raise RuntimeError("call precompute() before solve()")
        with torch.cuda.device(self._dev_index):
            F32, BC, B = self._canon_rhs(f, bc)
            self.f, self.bc = F32.to(self.precision), BC.to(self.precision)
            st = self._stream()
            lib, ctx = self._lib, self._ctx
            n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
            exact = self._exact(n_loc) and B >= _MIN_B_EXACT and self._xop_dh is not None
            if exact:
                self._reserve(n_loc, B)
                self._ensure_recon_planes()
            torch.cuda.synchronize(self._dev_index)
            t0 = time.perf_counter()

            # projected right-hand sides and (last row) the boundary sums: one buffer, one all-reduce
            RS = torch.zeros((FP + 1, B), dtype=torch.float64, device=self._tdev)
            R = RS[:FP]
            if n_loc:
                if exact and self._gplanes is not None:
                    _lib.check(lib.fps_project_rhs_x(ctx, self._xop_dh.data_ptr(), self._gplanes.data_ptr(), n_loc,
                                                     F32.data_ptr(), F32.stride(0), B, R.data_ptr(), B, st),
                               "fps_project_rhs_x")                                                            # :301
                else:
                    _lib.check(self._fn("fps_project_rhs")(ctx, self._DHp.data_ptr(), n_loc, FP, F32.data_ptr(),
                                                         F32.stride(0), B, R.data_ptr(), B, st), "fps_project_rhs")
            RS[FP] = BC.to(torch.float64).sum(dim=0)                                                             # :302
            if self.distributed:
                self._allreduce(RS)
            sum_bc = RS[FP]
            U = torch.empty((n_loc + nbc_loc, B), dtype=self._fdtype, device=self._tdev)
            Fp = torch.empty((n_loc, B), dtype=self._fdtype, device=self._tdev)
            W = torch.empty((FP, B), dtype=torch.float64, device=self._tdev)
            bias = torch.empty(B, dtype=torch.float64, device=self._tdev)
            best = None
            for i, lam in enumerate(self.lambdas_pde):
                _lib.check(lib.fps_solve_factored(ctx, self._L64[i].data_ptr(), R.data_ptr(), B, B,
                                                  float(lam) / self.NO, self._s64.data_ptr(), sum_bc.data_ptr(),
                                                  self.NdO, W.data_ptr(), bias.data_ptr(), st),
                           "fps_solve_factored")                                                                # :305-306
                self._reconstruct(W, bias, B, U, Fp, st, exact)                                                  # :327-331
                if self.n_lambdas > 1:
                    best = self._select_lambda(i, best, U, Fp, W, bias, F32, BC, B, st)
            if self.n_lambdas > 1:
                _, U, Fp, W, bias, which = best
                self.lambda_index = which
                self.lambda_pde = self.lambdas_pde[int(which[0].item())]
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

    def _select_lambda(self, i, best, U, Fp, W, bias, F32, BC, B, st):
        """Loss sweep of solver.py:309-325 per right-hand side: mean squared PDE residual + mean squared boundary
        residual (fps_sse_columns), smallest total wins."""
        n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
        sse = torch.zeros((2, B), dtype=torch.float64, device=self._tdev)
        fn = self._fn("fps_sse_columns")
        if n_loc:
            _lib.check(fn(self._ctx, Fp.data_ptr(), Fp.stride(0), F32.data_ptr(), F32.stride(0), None, n_loc, B,
                          sse[0].data_ptr(), st), "fps_sse_columns(pde)")
        if nbc_loc:
            Ub = U[n_loc:]
            if BC.stride(0) == 0:      # B constants
                bcc = BC[0].to(torch.float64).contiguous()
                _lib.check(fn(self._ctx, Ub.data_ptr(), Ub.stride(0), None, 0, bcc.data_ptr(), nbc_loc, B,
                              sse[1].data_ptr(), st), "fps_sse_columns(bc)")
            else:
                BCc = BC.to(self._fdtype).contiguous()
                _lib.check(fn(self._ctx, Ub.data_ptr(), Ub.stride(0), BCc.data_ptr(), BCc.stride(0), None, nbc_loc, B,
                              sse[1].data_ptr(), st), "fps_sse_columns(bc)")
        cnt = torch.tensor([[float(n_loc)], [float(nbc_loc)]], dtype=torch.float64, device=self._tdev)
        if self.distributed:
            both = torch.cat([sse, cnt.expand(2, 1)], dim=1)
            self._allreduce(both)
            sse, cnt = both[:, :B], both[:, B:]
        loss = sse / cnt
        first = loss[:, 0].tolist()
        self.Ls_pde[i], self.Ls_bc[i] = first
        self.Ls[i] = first[0] + first[1]
        tot = loss[0] + loss[1]
        if best is None:
            return (tot, U.clone(), Fp.clone(), W.clone(), bias.clone(), torch.zeros(B, dtype=torch.long, device=self._tdev))
        take = tot < best[0]
        best[0][take] = tot[take]
        best[1][:, take] = U[:, take]
        best[2][:, take] = Fp[:, take]
        best[3][:, take] = W[:, take]
        best[4][take] = bias[take]
        best[5][take] = i
        return best

    def evaluate(self, x, y):
        """Extension: the last solution u(x, y) = H(x, y) w + b at arbitrary points of [0,1]^2, (n, B)."""
        if getattr(self, "_W64", None) is None:
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

    def _reconstruct(self, W, bias, B, U, Fp, st, exact=False):
        n_loc, nbc_loc = self._DHp.shape[0], self._Hbcp.shape[0]
        lib, ctx = self._lib, self._ctx
        if n_loc:
            if exact and isinstance(self._rplanes_h, torch.Tensor):
                _lib.check(lib.fps_reconstruct_x(ctx, self._xop_h.data_ptr(), self._rplanes_h.data_ptr(), n_loc, W.data_ptr(),
                                                 B, bias.data_ptr(), B, U.data_ptr(), B, st), "fps_reconstruct_x(u_pde)")
                _lib.check(lib.fps_reconstruct_x(ctx, self._xop_dh.data_ptr(), self._rplanes_dh.data_ptr(), n_loc,
                                                 W.data_ptr(), B, None, B, Fp.data_ptr(), B, st), "fps_reconstruct_x(f)")
            else:
                _lib.check(self._fn("fps_reconstruct")(ctx, self._Hp.data_ptr(), n_loc, FP, W.data_ptr(), B, bias.data_ptr(), B,
                                                       U.data_ptr(), B, st), "fps_reconstruct(u_pde)")
                _lib.check(self._fn("fps_reconstruct")(ctx, self._DHp.data_ptr(), n_loc, FP, W.data_ptr(), B, None, B,
                                                       Fp.data_ptr(), B, st), "fps_reconstruct(f)")
        if nbc_loc:
            _lib.check(self._fn("fps_reconstruct")(ctx, self._Hbcp.data_ptr(), nbc_loc, FP, W.data_ptr(), B, bias.data_ptr(), B,
                                                   U[n_loc:].data_ptr(), B, st), "fps_reconstruct(u_bc)")
