"""CPU-only: the C-ABI library loads and exports every symbol include/osl_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from common import ROOT, pkg


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "osl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(osl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    P = pkg()
    L = ctypes.CDLL(P.capi.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libosl_b200.so does not export %s" % s
    assert sorted(P.capi.EXPORTS) == syms, "capi.EXPORTS is out of sync with include/osl_b200.h"


def test_status_strings_and_version_need_no_gpu():
    P = pkg()
    assert P.lib().osl_status_string(0) == b"ok"
    assert P.lib().osl_status_string(-4).startswith(b"node pool overflow")
    assert b"osl_b200" in P.lib().osl_version()


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    P = pkg()
    h = ctypes.c_void_p()
    c = (ctypes.c_float * 3)(0, 0, 0)
    assert P.lib().osl_svo_create(ctypes.byref(h), c, 1.0, 0, 0, 0) == -1      # max_depth < 1
    assert P.lib().osl_svo_create(ctypes.byref(h), c, 1.0, 21, 0, 0) == -1     # max_depth > 20
    assert P.lib().osl_svo_create(ctypes.byref(h), c, -1.0, 8, 0, 0) == -1     # half_edge <= 0
    assert P.lib().osl_integrate_points(None, None, None, 3, None) == -1
