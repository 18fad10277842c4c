"""CPU: the C-ABI shared library loads and exports every symbol include/sypha_b200.h declares."""
import ctypes
import re

import pytest

from conftest import REPO
from sypha_b200 import _lib


def declared_symbols():
    text = (REPO / "include" / "sypha_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_table_matches_header():
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    lib = _lib.load()
    assert lib.sb200_version() == 100


def test_struct_layouts_match_c():
    """sizes the C compiler produces for the ABI structs (x86-64 SysV)."""
    assert ctypes.sizeof(_lib.sb200_caps) == 16
    assert ctypes.sizeof(_lib.sb200_params) == 104
    assert ctypes.sizeof(_lib.sb200_result) == 152


def test_struct_layouts_against_the_c_compiler(tmp_path):
    """the same sizes, asked of gcc compiling include/sypha_b200.h as C."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("needs gcc")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "sypha_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(sb200_caps),sizeof(sb200_params),sizeof(sb200_result),sizeof(sb200_node_delta),'
                   'sizeof(sb200_heur_result),sizeof(sb200_scp_model),sizeof(sb200_row_model));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-I", str(_lib.LIB_PATH.parents[2] / "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(t) for t in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(getattr(_lib, n)) for n in ("sb200_caps", "sb200_params", "sb200_result", "sb200_node_delta",
                                                      "sb200_heur_result", "sb200_scp_model", "sb200_row_model")]
    assert sizes == mine


def test_no_cpu_fallback_without_gpu():
    lib = _lib.load()
    if lib.sb200_device_count() > 0:
        pytest.skip("GPU present")
    import sypha_b200 as sb
    ws = sb.IpmWorkspace()
    with pytest.raises(sb.Sb200Error):
        sb.initializeIpmWorkspace(ws)
    from conftest import load_golden, node_from_instance
    inst, _ = load_golden("demo00")
    with pytest.raises(sb.Sb200Error):
        sb.solver_sparse_mehrotra_run(node_from_instance(inst), sb.SolverExecutionConfig())


def test_product_never_imports_oracle():
    for p in (REPO / "sypha_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            assert "oracle" not in p.read_text(), p
