"""The C-ABI library loads and exports every symbol include/dmrgx.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dmrgx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmrgx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("dmrgx_hshell_apply", "dmrgx_hshell_apply_host", "dmrgx_hshell_create", "dmrgx_eigs_smallest", "dmrgx_truncate",
                 "dmrgx_rotate", "dmrgx_block_set_operator", "dmrgx_kron_create", "dmrgx_block_enlarge", "dmrgx_expect"):
        assert must in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "dmrg.x_b200", "libdmrgx_b200.so")
    assert os.path.exists(path), "build with __graft_entry__.build()"
    lib = C.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_every_header_entry_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "dmrgx.h")).read()
    assert len(re.findall(r"(src|include)/[A-Za-z]+\.(cpp|hpp):\d+", src)) >= 20


def test_product_does_not_touch_the_oracle():
    """The product path never imports, links or executes anything under oracle/."""
    pkg = os.path.join(ROOT, "dmrg.x_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "Makefile" and "oracle" not in txt, os.path.join(dp, f)
