import os
import pathlib
import sys

# Several ranks of a group share the one GPU of the test box (tests/test_comm.py): their kernels wait for each other's
# flags, and CUDA's lazy module loading can stall the first launch of a kernel behind such a waiting kernel.  Must be set
# before CUDA is initialised (csrc/cab_comm.cu connect_blobs refuses such a group otherwise).
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # two streams per rank, up to eight ranks on the GPU

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import pkgpath  # noqa: E402

pkgpath.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle

    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def kat():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "shape_data_kat.npz")
