"""The C-ABI library loads and exports every symbol include/b200caps.h declares (no compute here)."""
import ctypes
import os
import re

from b200caps import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200caps.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2c_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    assert os.path.isfile(_abi.LIB_PATH), "libb200caps.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_abi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    assert declared_symbols() == _abi.EXPORTS


def test_struct_sizes_match_header_layout(tmp_path):
    """sizeof/offsetof from the real header (gcc) == the ctypes mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "b200caps.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(b2c_conv_class), sizeof(b2c_conv_desc), '
        'sizeof(b2c_wgrad_desc), offsetof(b2c_conv_desc, cls), offsetof(b2c_conv_desc, bn_tile), '
        'offsetof(b2c_wgrad_desc, atomic));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_abi.ConvClass), ctypes.sizeof(_abi.ConvDesc), ctypes.sizeof(_abi.WgradDesc),
            _abi.ConvDesc.cls.offset, _abi.ConvDesc.bn_tile.offset, _abi.WgradDesc.atomic.offset]
    assert got == want, (got, want)


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", "/nonexistent/libb200caps.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _abi.lib()


def test_invalid_arguments_return_error_without_gpu():
    d = _abi.ConvDesc()
    with __import__("pytest").raises(RuntimeError, match="null tensor|conv_fprop"):
        _abi.call("b2c_conv_fprop", ctypes.byref(d), None)
    assert _abi.lib().b2c_version() == 100
