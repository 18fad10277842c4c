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
    assert ctypes.sizeof(_lib.sb200_params) == 88
    assert ctypes.sizeof(_lib.sb200_result) == 144


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
