import pathlib
import sys

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
