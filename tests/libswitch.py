"""Test-only hook: re-point the ctypes binding at another build of the C ABI (the host-logic emulation library
tests/plancheck/libdmrgx_plancheck.so on machines without a GPU).  Lives under tests/ on purpose: the product binding
(dmrg.x_b200/__init__.py) loads libdmrgx_b200.so and nothing else."""


def use_library(P, path):
    """path=None restores the product library."""
    import os
    P._lib = None
    P.LIB_PATH = path or os.path.join(os.path.dirname(os.path.abspath(P.__file__)), "libdmrgx_b200.so")
