"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/cloud_algos_b200.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes
import pathlib
import re

import pytest

from mapping_private_b200 import cab

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built():
    cab.build()
    return cab.LIB_PATH


def test_header_symbols_exported(built):
    header = (ROOT / "include" / "cloud_algos_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(cab_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(cab.EXPORTS)
    lib = ctypes.CDLL(str(built))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cab_version() >= 100


def test_no_cpu_fallback(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cab.CabError, match="no CUDA device|CUDA"):
        cab.Context(0)


def test_sass_is_sm100a(built):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", str(built)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)
