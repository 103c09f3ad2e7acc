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
    # map growth and camera tracking entry points
    assert P.lib().osl_svo_expand(None, 1) == -1
    assert P.lib().osl_svo_max_depth(None) == 0
    assert P.lib().osl_integrate_depth_posed(None, None, None, 4, 4, 1.0, 1.0, None, None) == -1
    t = ctypes.c_void_p()
    assert P.lib().osl_tracker_create(None, 640, 480, 500.0, 500.0, 0, 0) == -1          # no out pointer
    assert P.lib().osl_tracker_create(ctypes.byref(t), 642, 480, 500.0, 500.0, 0, 0) == -1  # width not a multiple of 4
    assert P.lib().osl_tracker_create(ctypes.byref(t), 640, 480, 0.0, 500.0, 0, 0) == -1  # focal length <= 0
    assert P.lib().osl_tracker_update(None, None, None) == -1
    assert P.lib().osl_tracker_update_host(None, None, None) == -1
    assert P.lib().osl_tracker_get_pose(None, None, None, None, None, None) == -1
    assert P.lib().osl_tracker_pose_device(None, None) == -1
    assert P.lib().osl_tracker_view(None, 0, None, None, None, None) == -1
    assert P.lib().osl_tracker_reset(None) == -1
    P.lib().osl_tracker_destroy(None)                                                   # a no-op
    A, b = (ctypes.c_float * 36)(), (ctypes.c_float * 6)()
    assert P.lib().osl_icp_cost(None, None, None, None, 10, 0, A, b, None, None) == -1
    assert P.lib().osl_bilateral_filter(None, None, 8, 8, None) == -1
    assert P.lib().osl_subsample_depth(None, None, 8, 8, None) == -1
    assert P.lib().osl_subsample_f32(None, None, 8, 8, None) == -1
    assert P.lib().osl_generate_normal_map(None, None, 8, 8, None) == -1
    assert P.lib().osl_transform_normal_map(None, None, 8, None) == -1
    assert P.lib().osl_color_to_intensity(None, None, 8, None) == -1
