"""CPU: host-side logic and the C-ABI surface (no compute calls - there is no GPU here)."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from fast_poisson_solver_b200 import _lib

    header = open(os.path.join(ROOT, "include", "fps_sm100.h")).read()
    declared = set(re.findall(r"\b(fps_[a-z0-9_]+)\s*\(", header))
    declared -= {"fps_ctx", "fps_weights"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fps_version() == 100


def test_ctypes_prototypes_match_header_arity():
    """Every prototype of include/fps_sm100.h and its ctypes twin in _lib.SIGNATURES take the same number of arguments."""
    from fast_poisson_solver_b200 import _lib

    header = open(os.path.join(ROOT, "include", "fps_sm100.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = dict(re.findall(r"\b(fps_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header))
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert name in protos, name
        args = protos[name].strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(argtypes), (name, n, len(argtypes))


def test_library_contains_blackwell_tensor_and_tma_instructions():
    """SASS evidence without a GPU: tcgen05 MMAs in fp16 (features) and int8 (exact Gram / projection /
    reconstruction), TMEM loads/stores and TMA loads are in the shipped library; no legacy HMMA path."""
    import subprocess
    from fast_poisson_solver_b200 import _lib

    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")
    assert "DMMA" in sass          # fp64 tensor path of the small / float64-mode kernels


def test_library_is_sm100a_only():
    import subprocess
    from fast_poisson_solver_b200 import _lib

    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_create_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fast_poisson_solver_b200 import Solver, _lib, weights

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Solver(use_weights=False)
    hw = weights.HostWeights(weights.default_init(0))
    ctx = ctypes.c_void_p()
    rc = _lib.load().fps_create(ctypes.byref(ctx), 0, ctypes.byref(hw.struct), 0, 0)
    assert rc != 0 and not ctx.value
    assert b"no CUDA device" in _lib.load().fps_last_error()


def test_default_init_matches_reference_digest(golden):
    from fast_poisson_solver_b200 import weights

    sd = weights.default_init(0)
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    assert h.hexdigest() == str(golden("features_seed0.npz")["weights_sha256"])


def test_state_dict_prefix_stripping_and_validation():
    from fast_poisson_solver_b200 import weights

    sd = weights.default_init(3)
    wrapped = {"module._orig_mod." + k: v for k, v in sd.items()}
    clean = weights.load_state_dict(wrapped)
    assert set(clean) == set(sd)
    weights.validate(clean)
    bad = dict(clean)
    bad["model.4.weight"] = torch.zeros(700, 699)
    with pytest.raises(ValueError, match="model.4.weight"):
        weights.validate(bad)
    del bad["model.4.weight"]
    with pytest.raises(KeyError):
        weights.validate(bad)


def test_format_input_contract():
    from fast_poisson_solver_b200 import format_input

    a, b, c, d = format_input([[1, 2, 3], np.arange(4.0).reshape(2, 2), torch.ones(5, dtype=torch.float64), 2.5],
                              torch.float32, "cpu")
    assert a.shape == (3, 1) and b.shape == (4, 1) and c.shape == (5, 1) and d == 2.5
    assert a.dtype == b.dtype == c.dtype == torch.float32
    (e,) = format_input([np.zeros((2, 3))], torch.float64, "cpu", reshape=False)
    assert e.shape == (2, 3) and e.dtype == torch.float64
    (f,) = format_input([[1.0, 2.0]], as_array=True)
    assert isinstance(f, np.ndarray) and f.shape == (2, 1)


def test_perlin_family_host_generator():
    """Host mirror of the device Perlin generator: deterministic, smooth across lattice cells, parameter draws in the
    ranges of source_functions.py:173-227."""
    from fast_poisson_solver_b200.sources import perlin_params, perlin_sources_numpy

    p = perlin_params(64, seed=1)
    assert p.shape == (64, 6, 3)
    used = p[:, :5, 1] != 0
    assert used.sum(1).min() >= 1 and used.sum(1).max() <= 5
    assert p[:, :5, 0][used].min() >= 0.1 and p[:, :5, 0][used].max() <= 12 and np.abs(p[:, :5, 1]).max() <= 100
    assert 0.5 < (p[:, 5, 1] != 0).mean() < 0.9          # the modulation layer is applied with probability 0.7
    x = np.linspace(0, 1, 2001)
    F = perlin_sources_numpy(x, np.full_like(x, 0.37), p[:4])
    assert np.array_equal(F, perlin_sources_numpy(x, np.full_like(x, 0.37), p[:4]))
    assert np.isfinite(F).all() and np.abs(F).max() > 0.1
    assert np.abs(np.diff(F, axis=0)).max() < 0.05 * np.abs(F).max() + 1.0   # continuous (quintic fade)


def test_shard_bounds_cover_everything():
    from fast_poisson_solver_b200.parallel import shard_bounds

    for n, w in ((10, 3), (4100, 8), (7, 8), (1 << 20, 4)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
