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


def test_header_compiles_as_plain_c_and_a_c_program_links_the_library(tmp_path):
    """The boundary is a C ABI: include/transgo_b200.h must be valid C99 (no C++-isms, no torch types), and a C translation unit
    that references entry points of every family must link against the shared library (nothing is CALLED here: no GPU)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from transgo_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)
    src = tmp_path / "consumer.c"
    src.write_text("""
#include <stdio.h>
#include "transgo_b200.h"
int main(int argc, char** argv) {
    /* take the addresses so that the linker must resolve them; call only what needs no device */
    void* fns[] = {(void*)Reset, (void*)Step, (void*)getLegalAction, (void*)Encode, (void*)getScore,
                   (void*)tg_env_step, (void*)tg_env_playout_scored, (void*)tg_mcts_wave_begin, (void*)tg_mcts_wave_end,
                   (void*)tg_traj_assemble, (void*)tg_net_attention_fused2, (void*)tg_net_conv3x3_head, (void*)tg_net_stem};
    printf("%d %d %d %d\\n", tg_version(), tg_state_bytes(9), tg_state_bytes(19), (int)(sizeof(fns) / sizeof(fns[0])));
    return argc > 100 ? (int)(size_t)fns[0] : 0;
}
""")
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libtransgo_b200.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) >= 100 and out[1:3] == ["48", "128"]
