"""Weights of the feature network: default initialisation, state_dict ingestion, host packing.

Reference behaviour being mirrored (paths relative to /root/reference/fast_poisson_solver/):
  * architecture: resources/infos.yaml:838-859 (width 700, depth 6, ffm num 200 scale 10, tanh,
    out 800, no batch-norm, dropout 0) -> PINN.__init__ model.py:23-67.
  * ``use_weights=False``: the network keeps torch's default initialisation drawn right after
    ``torch.manual_seed(seed)`` (solver.py:86-88, :139).
  * ``use_weights=True``: ``resources/final.pt`` is loaded and the 'module.' / '_orig_mod.' key
    prefixes are stripped (solver.py:141-149).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib

MODEL_INFO = {
    "activation_fn": "tanh",
    "depth": 6,
    "dropout": 0,
    "ffm": {"activation_fn": "tanh", "num": 200, "num_inputs": 2, "num_outputs": 700, "req_grad": False, "scale": 10},
    "network": [700] * 6,
    "out": 800,
    "use_bn": False,
    "weights_init": None,
    "width": 700,
}
HIDDEN_KEYS = ["model.1", "model.4", "model.7", "model.10", "model.13"]  # Sequential indices, model.py:49-60
EXPECTED_SHAPES = {
    "mapping.freqs": (2, 200),
    "mapping.coefficients": (200,),
    "mapping.biases": (200,),
    "mapping.linear.weight": (700, 400),
    "mapping.linear.bias": (700,),
    **{k + ".weight": (700, 700) for k in HIDDEN_KEYS},
    **{k + ".bias": (700,) for k in HIDDEN_KEYS},
}
RESOURCE_WEIGHTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources", "final.pt")


def default_init(seed: int) -> dict:
    """The state_dict the reference ends up with for ``use_weights=False, seed=seed``.

    Draw order follows PINN.__init__: mapping.freqs = randn(2,200)/2*10 (model.py:124), ones,
    zeros, mapping.linear, the five hidden Linear layers, the (unused) 800-wide output head.
    The torch global RNG is left seeded exactly as the reference leaves it.
    """
    torch.manual_seed(seed)
    sd = {
        "mapping.freqs": torch.randn(2, 200) / 2 * 10.0,
        "mapping.coefficients": torch.ones(200),
        "mapping.biases": torch.zeros(200),
    }
    for key, (fan_in, fan_out) in [("mapping.linear", (400, 700))] + [(k, (700, 700)) for k in HIDDEN_KEYS] + [
        ("out", (700, 800))
    ]:
        lin = torch.nn.Linear(fan_in, fan_out)
        sd[key + ".weight"] = lin.weight.detach()
        sd[key + ".bias"] = lin.bias.detach()
    return sd


def load_state_dict(source) -> dict:
    """Accept a path / file object / mapping; strip DataParallel and torch.compile key prefixes."""
    if isinstance(source, (str, os.PathLike)) or hasattr(source, "read"):
        source = torch.load(source, map_location="cpu")
    sd = {}
    for k, v in source.items():
        k = k.replace("module.", "").replace("_orig_mod.", "")
        sd[k] = torch.as_tensor(v).detach().to("cpu")
    return sd


def validate(sd: dict) -> None:
    for k, shape in EXPECTED_SHAPES.items():
        if k not in sd:
            raise KeyError(f"state_dict is missing '{k}' (expected the PINN of resources/infos.yaml: width 700, depth 6, ffm 200)")
        if tuple(sd[k].shape) != shape:
            raise ValueError(f"state_dict['{k}'] has shape {tuple(sd[k].shape)}, expected {shape}; "
                             "only the shipped architecture is compiled into libfps_sm100")


class HostWeights:
    """float32 C-contiguous host copies + the ``fps_weights`` struct pointing at them."""

    def __init__(self, sd: dict):
        validate(sd)
        self.arrays = {k: np.ascontiguousarray(sd[k].to(torch.float32).numpy()) for k in EXPECTED_SHAPES}
        p = lambda k: self.arrays[k].ctypes.data_as(C.POINTER(C.c_float))
        self.struct = _lib.FpsWeights()
        self.struct.h_freqs = p("mapping.freqs")
        self.struct.h_coefficients = p("mapping.coefficients")
        self.struct.h_biases = p("mapping.biases")
        self.struct.h_w0 = p("mapping.linear.weight")
        self.struct.h_b0 = p("mapping.linear.bias")
        for i, k in enumerate(HIDDEN_KEYS):
            self.struct.h_wh[i] = p(k + ".weight")
            self.struct.h_bh[i] = p(k + ".bias")

    def sha256(self) -> str:
        """Identity of the weights the feature matrices were computed with (stored in saved precompute state)."""
        import hashlib

        h = hashlib.sha256()
        for k in sorted(self.arrays):
            h.update(k.encode())
            h.update(self.arrays[k].tobytes())
        return h.hexdigest()
