import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffqcqp_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):  # a checkout without the built artefact: compile it (nvcc is in the image);
        build.build()                       # the product path itself never builds or falls back, it raises
    return _lib.load()  # raises (test error, not skip) if the extension is missing
