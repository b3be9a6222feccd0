"""Registers the package directory ``dmrg.x_b200/`` (not a valid Python identifier) as module ``dmrgx_b200``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def load_package():
    if "dmrgx_b200" in sys.modules:
        return sys.modules["dmrgx_b200"]
    path = os.path.join(ROOT, "dmrg.x_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("dmrgx_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dmrgx_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
