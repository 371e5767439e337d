"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/dsb200.h declares,
and fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from deepsignal_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dsb200.h")).read()
    return sorted(set(re.findall(r"DSB_API\s+[\w\s\*]+?\b(dsb_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.dsb_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.dsb_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.dsb_last_error(None) or b"CUDA" in lib.dsb_last_error(None)
    from deepsignal_b200.engine import Engine
    with pytest.raises(_lib.DsbError):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deepsignal_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn
