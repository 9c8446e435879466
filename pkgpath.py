"""Makes the hyphenated package directory `mapping-private_b200/` importable as
`mapping_private_b200` (used by tests, bench.py and __graft_entry__.py)."""
import importlib.util
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent
PKG_DIR = ROOT / "mapping-private_b200"
PKG_NAME = "mapping_private_b200"


def load():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(
        PKG_NAME, PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)]
    )
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
