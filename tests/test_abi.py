"""CPU tests of the drop-in boundary: the library loads and exports every symbol include/transgo_b200.h declares
(no compute is attempted without a GPU), and the host side refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "transgo_b200.h")

REFERENCE_SYMBOLS = ["Init", "Reset", "Step", "Step_", "checkAction", "isTerminated", "Encode", "getScore",
                     "getTerritory", "getLegalAction", "getLegalNoEye", "Show", "getPlayer", "getStep", "getSubEncode"]


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)          # preprocessor lines
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(names))


def test_header_declares_the_reference_abi():
    names = declared_symbols()
    for s in REFERENCE_SYMBOLS:           # GoEnv/cpp_src/go_env.h:23-71
        assert s in names, s
    assert any(n.startswith("tg_mcts_") for n in names)
    assert "tg_env_step" in names and "tg_env_encode" in names


def test_library_exports_every_declared_symbol():
    from transgo_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert L.tg_version() >= 100
    L.tg_state_bytes.argtypes = [ctypes.c_int]
    assert L.tg_state_bytes(9) == 48 and L.tg_state_bytes(19) == 128
    assert L.tg_legal_words(9) == 4 and L.tg_legal_words(19) == 12


def test_no_cpu_fallback():
    import torch
    from transgo_b200 import _lib
    from transgo_b200.environment import BatchedGoEnv
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.TransgoLibraryError):
        BatchedGoEnv(4)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under transgo_b200/ may reference it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "transgo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle_go|oracle/_ref", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad
