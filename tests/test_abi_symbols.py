"""The C-ABI library loads and exports every symbol include/hemelb_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from hemelb_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hemelb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hlb_(?:gpu|dom|xtr|part)_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(capi.LIB_PATH):
        from hemelb_b200 import build
        build.build()
    L = ctypes.CDLL(capi.LIB_PATH)
    for s in _declared():
        assert hasattr(L, s), s


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to construct (and never routes to the oracle)."""
    import ctypes as C
    L = capi.lib()
    n = C.c_int(0)
    rc = L.hlb_gpu_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    cfg = capi.HlbConfig()
    cfg.lattice, cfg.tau = 19, 0.8
    h = C.c_void_p()
    assert L.hlb_gpu_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CPU fallback" in L.hlb_gpu_last_error()
    for f in os.listdir(os.path.join(ROOT, "hemelb_b200")):
        if f.endswith(".py"):
            assert "import oracle" not in open(os.path.join(ROOT, "hemelb_b200", f)).read()
