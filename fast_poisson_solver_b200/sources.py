"""Synthetic source functions generated on the device (the step *before* the hot path).

`sine_params` draws the parameters of the reference's 'sin' family (data_utils/source_functions.py:24-73:
1-9 terms, k ~ U(-10,10), o ~ U(-1,1), s ~ U(-100,100), second factor sin or cos); `sine_sources`
evaluates B such sources on the solver's own collocation points straight into an (N, B) device
matrix, and `sine_sources_numpy` is the same formula on the host for checks and for the FD oracle.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

MAX_TERMS = 9


def sine_params(B: int, seed: int = 0) -> np.ndarray:
    """(B, 9, 6) float64: {k1, o1, k2, o2, s, use_cos}; unused terms have s = 0."""
    rng = np.random.default_rng(seed)
    p = np.zeros((B, MAX_TERMS, 6))
    for j in range(B):
        t = int(rng.integers(1, MAX_TERMS + 1))
        p[j, :t, 0] = rng.uniform(-10, 10, t)
        p[j, :t, 1] = rng.uniform(-1, 1, t)
        p[j, :t, 2] = rng.uniform(-10, 10, t)
        p[j, :t, 3] = rng.uniform(-1, 1, t)
        p[j, :t, 4] = rng.uniform(-100, 100, t)
        p[j, :t, 5] = rng.integers(0, 2, t)
    return p


def sine_sources_numpy(x, y, params) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1, 1)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1, 1)
    k1, o1, k2, o2, s, uc = (params[None, :, :, i] for i in range(6))
    second = np.where(uc != 0, np.cos(np.pi * k2 * (y - o2)), np.sin(np.pi * k2 * (y - o2)))
    return (s * np.sin(np.pi * k1 * (x - o1)) * second).sum(axis=2)


def sine_sources(solver, params, x=None, y=None, out=None) -> torch.Tensor:
    """Evaluate the sources on the solver's PDE points (or on given device float64 coordinates)."""
    dev = solver._tdev
    x64 = solver._xy64[0] if x is None else x
    y64 = solver._xy64[1] if y is None else y
    n, (B, T, _) = x64.numel(), params.shape
    pd = torch.as_tensor(np.ascontiguousarray(params), dtype=torch.float64, device=dev)
    if out is None:
        out = torch.empty((n, B), dtype=solver._fdtype, device=dev)
    fn = solver._lib.fps_sine_sources_f64 if out.dtype == torch.float64 else solver._lib.fps_sine_sources
    _lib.check(fn(solver._ctx, x64.data_ptr(), y64.data_ptr(), n, pd.data_ptr(), T, B, out.data_ptr(), out.stride(0),
                  solver._stream()), "fps_sine_sources")
    return out
