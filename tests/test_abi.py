"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol that
include/b200sv.h declares, and the ctypes table in qrack_b200/_abi.py covers exactly that set.  No compute calls."""
import ctypes
import os
import re

from qrack_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200sv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(b200sv_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_exported_and_bound():
    declared = _declared()
    assert len(declared) > 40
    _abi.build()
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libb200sv.so does not export %s" % name
    bound = set(_abi.SIGNATURES) | {"b200sv_last_error"}
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))


def test_abi_version_and_error_string():
    lib = _abi.load()
    assert lib.b200sv_abi_version() == 1
    assert isinstance(lib.b200sv_last_error(), bytes)


def test_no_cpu_fallback_without_library(tmp_path, monkeypatch):
    """The product path must fail loudly when the CUDA extension is missing."""
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "missing.so"))
    import pytest
    with pytest.raises(RuntimeError):
        _abi.load()


def test_reference_c_api_links_against_the_dropin():
    """dropin/Makefile also links the reference's own C API (src/pinvoke_api.cpp — what PyQrack and the other language
    bindings dlopen) on top of the drop-in QEngineCUDA: the shared object must load and export the binding surface."""
    import ctypes
    import os
    import pytest
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin", "_build", "f32", "libqrack_pinvoke.so")
    if not os.path.exists(so):
        pytest.skip("dropin/_build not built (needs /root/reference)")
    lib = ctypes.CDLL(so)
    for name in ("init_count_type", "destroy", "seed", "H", "MCX", "MCMtrx", "Prob", "M", "MAll", "Compose", "Decompose", "QFT", "ADD", "MUL"):
        assert hasattr(lib, name), name
