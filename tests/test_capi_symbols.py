"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/xyst_b200.h declares, and compute calls fail loudly without a device."""
import os
import re
import ctypes
import pytest
import xyst_b200
from xyst_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(xyst_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = xyst_b200.lib()
    names = declared("xyst_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert sorted(capi.SYMBOLS) == names


def test_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(xyst_b200.XystError, match="no CUDA device"):
        xyst_b200.Context()
