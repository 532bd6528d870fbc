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


# ---- Perlin family (data_utils/source_functions.py:173-227) ------------------------------------------------
# The reference sums 1-5 layers of the `perlin_noise` package's gradient noise (octaves ~ U(0.1,12), amplitude
# ~ U(-100,100), one integer seed per layer) and, with probability 0.7, multiplies the sum by one more layer
# (octaves ~ U(0.1,3), amplitude ~ U(-1,1)).  That package is a per-point Python loop and is absent here; sources are
# INPUTS to the path, so the generator below keeps the family (same parameter draws, same quintic fade, same
# lattice-gradient construction, coordinates scaled by `octaves`) but picks its lattice gradients (64 unit directions)
# with an integer hash, which the device kernel (csrc/sources.cu: perlin_sources_kernel) evaluates identically.
PERLIN_LAYERS = 5          # rows 0..4 of a parameter block: additive layers {octave, amplitude, seed}
PERLIN_ROWS = 6            # row 5: multiplicative modulation layer (amplitude 0 = none)


def perlin_params(B: int, seed: int = 0) -> np.ndarray:
    """(B, 6, 3) float64: rows 0..4 additive layers {octave, amplitude, seed} (amplitude 0 = unused), row 5 the
    modulation layer of source_functions.py:213-220 (amplitude 0 = not applied)."""
    rng = np.random.default_rng(seed)
    p = np.zeros((B, PERLIN_ROWS, 3))
    for j in range(B):
        n = int(rng.integers(1, PERLIN_LAYERS + 1))
        p[j, :n, 0] = rng.uniform(0.1, 12.0, n)
        p[j, :n, 1] = rng.uniform(-100.0, 100.0, n)
        p[j, :n, 2] = rng.integers(1, 1000000, n)
        if rng.random() > 0.3:
            p[j, 5] = (rng.uniform(0.1, 3.0), rng.uniform(-1.0, 1.0), rng.integers(1, 1000000))
    return p


_GRAD = np.stack([np.cos(2.0 * np.pi * np.arange(64) / 64.0), np.sin(2.0 * np.pi * np.arange(64) / 64.0)], axis=1)


def _lattice_hash(ix, iy, seed):
    """32-bit integer mix of a lattice point and a layer seed (uint32 wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        h = (ix.astype(np.uint32) * np.uint32(0x9E3779B1)) ^ (iy.astype(np.uint32) * np.uint32(0x85EBCA77)) \
            ^ (np.uint32(seed) * np.uint32(0xC2B2AE3D))
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x7FEB352D)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x846CA68B)
        h ^= h >> np.uint32(16)
    return h


def _perlin_layer(x, y, octave, seed):
    px, py = x * octave, y * octave
    fx0, fy0 = np.floor(px), np.floor(py)
    ix, iy = fx0.astype(np.int64), fy0.astype(np.int64)
    fx, fy = px - fx0, py - fy0
    sx = fx * fx * fx * (fx * (fx * 6.0 - 15.0) + 10.0)
    sy = fy * fy * fy * (fy * (fy * 6.0 - 15.0) + 10.0)

    def corner(dx, dy):
        g = (_lattice_hash(ix + dx, iy + dy, int(seed)) >> np.uint32(26)).astype(np.int64)   # one of 64 directions
        return _GRAD[g, 0] * (fx - dx) + _GRAD[g, 1] * (fy - dy)

    n00, n10, n01, n11 = corner(0, 0), corner(1, 0), corner(0, 1), corner(1, 1)
    return (n00 * (1.0 - sx) + n10 * sx) * (1.0 - sy) + (n01 * (1.0 - sx) + n11 * sx) * sy


def perlin_sources_numpy(x, y, params) -> np.ndarray:
    """(n, B) float64: host mirror of fps_perlin_sources."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    out = np.zeros((x.size, params.shape[0]))
    for j, blk in enumerate(params):
        u = np.zeros_like(x)
        for octave, amp, seed in blk[:PERLIN_LAYERS]:
            if amp != 0.0:
                u += amp * _perlin_layer(x, y, octave, seed)
        octave, amp, seed = blk[5]
        if amp != 0.0:
            u *= amp * _perlin_layer(x, y, octave, seed)
        out[:, j] = u
    return out


def perlin_sources(solver, params, x=None, y=None, out=None) -> torch.Tensor:
    """Evaluate B Perlin-family sources on the solver's PDE points (or on given device float64 coordinates) on the
    device (csrc/sources.cu: perlin_sources_kernel); (n, B) in the solver's feature dtype."""
    dev = solver._tdev
    x64 = solver._xy64[0] if x is None else x
    y64 = solver._xy64[1] if y is None else y
    n, B = x64.numel(), params.shape[0]
    assert params.shape[1:] == (PERLIN_ROWS, 3)
    pd = torch.as_tensor(np.ascontiguousarray(params), dtype=torch.float64, device=dev)
    if out is None:
        out = torch.empty((n, B), dtype=solver._fdtype, device=dev)
    fn = solver._lib.fps_perlin_sources_f64 if out.dtype == torch.float64 else solver._lib.fps_perlin_sources
    _lib.check(fn(solver._ctx, x64.data_ptr(), y64.data_ptr(), n, pd.data_ptr(), B, out.data_ptr(), out.stride(0),
                  solver._stream()), "fps_perlin_sources")
    return out


def grid_points(solver, G: int, rank: int = 0, world: int = 1):
    """Device float64 coordinates of the reference generator's geometry (data.py:192-217): this rank's block of the
    G x G interior grid and of the 4G + 4 boundary ring (block sharding as parallel.shard_bounds)."""
    from .parallel import shard_bounds

    dev = solver._tdev
    lo, hi = shard_bounds(G * G, rank, world)
    blo, bhi = shard_bounds(4 * G + 4, rank, world)
    x = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    y = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    xb = torch.empty(bhi - blo, dtype=torch.float64, device=dev)
    yb = torch.empty(bhi - blo, dtype=torch.float64, device=dev)
    _lib.check(solver._lib.fps_grid_points(solver._ctx, G, lo, hi - lo, x.data_ptr(), y.data_ptr(), blo, bhi - blo,
                                           xb.data_ptr(), yb.data_ptr(), solver._stream()), "fps_grid_points")
    return x, y, xb, yb
