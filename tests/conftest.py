import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    return load


@pytest.fixture(scope="session")
def weights0():
    from oracle import fps_oracle as o

    return o.init_weights(0)


@pytest.fixture(scope="session", autouse=True)
def _chdir_tmp(tmp_path_factory):
    """Solver() creates ./precomputed-resources in the CWD (as the reference does); keep it out of the repo."""
    d = tmp_path_factory.mktemp("cwd")
    old = os.getcwd()
    os.chdir(d)
    yield
    os.chdir(old)
