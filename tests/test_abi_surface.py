"""CPU: the C-ABI library loads without a GPU, exports every symbol include/bsmr_b200.h declares,
and refuses to compute (no CPU fallback)."""
import ctypes
import os
import re

import pytest


def declared_symbols(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsmr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    names = declared_symbols(pkg.HEADER)
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libbsmr_b200.so does not export " + n


def test_binding_covers_header(pkg):
    L = pkg.lib()
    for n in declared_symbols(pkg.HEADER):
        getattr(L, n)


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.BsmrError) as e:
        pkg.Context(0)
    assert e.value.status == 2  # BSMR_ERR_NO_DEVICE
    assert "no CPU" in str(e.value) or "CPU" in str(e.value)


def test_product_does_not_touch_oracle(pkg):
    """The product sources must not reference oracle/ (parity claims depend on it)."""
    root = os.path.dirname(pkg.HERE)
    for d, _, files in os.walk(pkg.HERE):
        if "_obj" in d or "__pycache__" in d:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp", ".py")):
                src = open(os.path.join(d, f), errors="replace").read()
                assert "bsmr_oracle" not in src and "oracle/" not in src and "libbsmr_ref" not in src, os.path.join(d, f)
    assert os.path.isdir(os.path.join(root, "oracle"))


def test_status_strings(pkg):
    L = pkg.lib()
    assert L.bsmr_status_string(0) == b"ok"
    assert b"CPU" in L.bsmr_status_string(2)
    assert b"bsmr_b200" in L.bsmr_version()
