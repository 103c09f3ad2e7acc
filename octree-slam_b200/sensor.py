"""Host-side Python mirror of the reference's sensor interface for camera tracking, over the C ABI.

  sensor::bilateralFilter / subsampleDepth / subsample / generateNormalMap / transformNormalMap / colorToIntensity
      (include/octree_slam/sensor/image_kernels.h, src/sensor/image_kernels.cu:104-321)
  sensor::computeICPCost2  (include/octree_slam/sensor/localization_kernels.h:40, localization_kernels.cu:313-330)
  sensor::RGBDCamera       (include/octree_slam/sensor/rgbd_camera.h:17-82, src/sensor/rgbd_camera.cpp:21-191)
torch only holds device buffers; every computation happens in libosl_b200.so.
"""
import ctypes as C

import numpy as np

from .capi import _check, _f, lib, mat_colmajor
from .world import _dev, _torch


def bilateralFilter(depth, device=0):
    torch = _torch()
    d = _dev(depth, np.uint16, device)
    h, w = d.shape
    out = torch.empty_like(d)
    _check(lib().osl_bilateral_filter(d.data_ptr(), out.data_ptr(), w, h, None), "osl_bilateral_filter")
    return out


def subsampleDepth(depth, device=0):
    torch = _torch()
    d = _dev(depth, np.uint16, device)
    h, w = d.shape
    out = torch.empty((h // 2, w // 2), dtype=torch.uint16, device=d.device)
    _check(lib().osl_subsample_depth(d.data_ptr(), out.data_ptr(), w, h, None), "osl_subsample_depth")
    return out


def subsample(img, device=0):
    torch = _torch()
    a = _dev(img, np.float32, device)
    h, w = a.shape
    out = torch.empty((h // 2, w // 2), dtype=torch.float32, device=a.device)
    _check(lib().osl_subsample_f32(a.data_ptr(), out.data_ptr(), w, h, None), "osl_subsample_f32")
    return out


def generateNormalMap(vertex_map, width, height):
    torch = _torch()
    out = torch.empty_like(vertex_map)
    _check(lib().osl_generate_normal_map(vertex_map.data_ptr(), out.data_ptr(), width, height, None),
           "osl_generate_normal_map")
    return out


def transformNormalMap(normal_map, trans):
    _check(lib().osl_transform_normal_map(normal_map.data_ptr(), _f(mat_colmajor(trans)), normal_map.shape[0], None),
           "osl_transform_normal_map")
    return normal_map


def colorToIntensity(rgb, device=0):
    torch = _torch()
    c = _dev(rgb, np.uint8, device).reshape(-1, 3)
    out = torch.empty(c.shape[0], dtype=torch.float32, device=c.device)
    _check(lib().osl_color_to_intensity(c.data_ptr(), out.data_ptr(), c.shape[0], None), "osl_color_to_intensity")
    return out


def computeICPCost2(last_vertex, last_normal, this_vertex, this_normal, exact_jacobian=False):
    """-> (A 6x6, b 6, pairs); the four maps are CUDA float32 tensors of shape (n, 3)"""
    A, b, pairs = (C.c_float * 36)(), (C.c_float * 6)(), C.c_int()
    _check(lib().osl_icp_cost(last_vertex.data_ptr(), last_normal.data_ptr(), this_vertex.data_ptr(),
                              this_normal.data_ptr(), last_vertex.shape[0], int(bool(exact_jacobian)), A, b,
                              C.byref(pairs), None), "osl_icp_cost")
    return np.array(A, dtype=np.float32).reshape(6, 6), np.array(b, dtype=np.float32), pairs.value


class RGBDCamera:
    """sensor::RGBDCamera: frame-to-frame ICP tracking.  exact_jacobian=False reproduces the reference (whose
    Jacobian and pose accumulation are broken, quirks Q17/Q18); True is the corrected tracker."""

    PYRAMID_DEPTH = 3
    PYRAMID_ITERS = (10, 5, 4)

    def __init__(self, width, height, focal_length, exact_jacobian=False, device=0):
        self.width_, self.height_ = int(width), int(height)
        self.focal_length_ = (float(focal_length[0]), float(focal_length[1]))
        self.device = device
        h = C.c_void_p()
        _check(lib().osl_tracker_create(C.byref(h), self.width_, self.height_, self.focal_length_[0],
                                        self.focal_length_[1], int(bool(exact_jacobian)), device),
               "osl_tracker_create")
        self._h = h
        self._keep = None

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().osl_tracker_destroy(self._h)
            self._h = None

    def reset(self):
        _check(lib().osl_tracker_reset(self._h), "osl_tracker_reset")

    def update(self, depth):
        """rgbd_camera.cpp:53-191 for one depth image (numpy (h, w) uint16 in host memory, or a CUDA tensor)"""
        if isinstance(depth, np.ndarray):
            d = np.ascontiguousarray(depth, dtype=np.uint16)
            assert d.shape == (self.height_, self.width_)
            self._keep = d  # the copy is asynchronous
            _check(lib().osl_tracker_update_host(self._h, d.ctypes.data_as(C.c_void_p), None),
                   "osl_tracker_update_host")
        else:
            assert depth.is_cuda and tuple(depth.shape) == (self.height_, self.width_)
            self._keep = depth
            _check(lib().osl_tracker_update(self._h, depth.data_ptr(), None), "osl_tracker_update")

    def pose_device(self):
        """device address of the pose matrix (osl_tracker_pose_device), for SVO.integrate_depth_tracked"""
        p = C.c_void_p()
        _check(lib().osl_tracker_pose_device(self._h, C.byref(p)), "osl_tracker_pose_device")
        return p.value

    def _get(self):
        pose, pos, ori = (C.c_float * 16)(), (C.c_float * 3)(), (C.c_float * 9)()
        lost, pairs = C.c_int(), C.c_int()
        _check(lib().osl_tracker_get_pose(self._h, pose, pos, ori, C.byref(lost), C.byref(pairs)),
               "osl_tracker_get_pose")
        return (np.array(pose, dtype=np.float32).reshape(4, 4).T, np.array(pos, dtype=np.float32),
                np.array(ori, dtype=np.float32).reshape(3, 3).T, bool(lost.value), pairs.value)

    def pose(self):
        """the matrix main.cpp:40 applies to the vertex map (math convention m[r][c])"""
        return self._get()[0]

    def position(self):
        return self._get()[1]

    def orientation(self):
        return self._get()[2]

    @property
    def lost(self):
        return self._get()[3]

    @property
    def pairs(self):
        return self._get()[4]

    def level(self, i):
        """(vertex, normal) maps of pyramid level i of the last frame as numpy arrays (n, 3)"""
        torch = _torch()
        self._get()
        pv, pn, w, h = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_int()
        _check(lib().osl_tracker_view(self._h, i, C.byref(pv), C.byref(pn), C.byref(w), C.byref(h)),
               "osl_tracker_view")
        n = w.value * h.value
        out = []
        for p in (pv, pn):
            t = torch.empty((n, 3), dtype=torch.float32, device="cuda:%d" % self.device)
            _check(lib().osl_copy_device(t.data_ptr(), p, n * 12), "osl_copy_device")
            out.append(t.cpu().numpy())
        return out[0], out[1]
