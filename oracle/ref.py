"""ctypes binding of oracle/_ref/libosl_ref{,64}.so: the REFERENCE's own CUDA sources behind a headless C shim
(oracle/ref_shim.cu).  TEST INFRASTRUCTURE ONLY; needs a GPU.  Built by `make -C oracle ref` where /root/reference
is mounted; the .so travels to the GPU box (git-ignored, not gpurun-ignored)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(patched64=False):
    return os.path.join(HERE, "_ref", "libosl_ref64.so" if patched64 else "libosl_ref.so")


def available(patched64=False):
    return os.path.exists(lib_path(patched64))


_libs = {}


def lib(patched64=False):
    if patched64 not in _libs:
        L = C.CDLL(lib_path(patched64))
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.ref_max_depth_from_resolution.restype = i32
        L.ref_max_depth_from_resolution.argtypes = [f32, f32]
        L.ref_create.restype = vp
        L.ref_create.argtypes = [fp, f32, i32]
        L.ref_destroy.argtypes = [vp]
        L.ref_size.argtypes = [vp]
        L.ref_download_pool.argtypes = [vp, vp]
        L.ref_upload_pool.argtypes = [vp, vp, i32]
        L.ref_integrate_depth_dev.argtypes = [vp, vp, vp, i32, i32, f32, f32, fp, dp]
        L.ref_integrate_depth.argtypes = [vp, vp, vp, i32, i32, f32, f32, fp, dp, dp]
        L.ref_vertex_map.argtypes = [vp, i32, i32, f32, f32, fp, vp]
        L.ref_bbox.argtypes = [vp, i32, fp]
        L.ref_integrate_points.argtypes = [vp, vp, vp, i32]
        L.ref_integrate_voxels.argtypes = [vp, vp, vp, i32]
        L.ref_extract_voxels.restype = C.c_longlong
        L.ref_extract_voxels.argtypes = [vp, i32, vp, vp, C.c_longlong]
        L.ref_raycast.argtypes = [vp, vp, i32, i32, f32, fp, dp]
        L.ref_bilateral.argtypes = [vp, i32, i32, vp]
        L.ref_subsample_depth.argtypes = [vp, i32, i32, vp]
        L.ref_subsample_f32.argtypes = [vp, i32, i32, vp]
        L.ref_normal_map.argtypes = [vp, i32, i32, vp]
        L.ref_transform_normals.argtypes = [vp, i32, fp]
        L.ref_color_to_intensity.argtypes = [vp, i32, vp]
        L.ref_icp_cost2.argtypes = [vp, vp, vp, vp, i32, i32, fp, fp]
        L.ref_tracker_create.restype = vp
        L.ref_tracker_create.argtypes = [i32, i32, f32, f32]
        L.ref_tracker_destroy.argtypes = [vp]
        L.ref_tracker_update.restype = C.c_double
        L.ref_tracker_update.argtypes = [vp, vp]
        L.ref_tracker_pose.argtypes = [vp, fp, fp]
        _libs[patched64] = L
    return _libs[patched64]


def _f(arr):
    arr = [float(x) for x in arr]
    return (C.c_float * len(arr))(*arr)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def mat_colmajor(m):
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T).reshape(16)


IDENTITY = np.eye(4, dtype=np.float32)


class RefSVO:
    """The reference's svo.cu pool driven exactly like Octree::addCloud / mainLoop drive it."""

    def __init__(self, center=(0, 0, 0), half_edge=1.0, max_depth=8, patched64=None):
        if patched64 is None:
            patched64 = max_depth > 10  # the unmodified reference is only correct for D <= 10 (svo.cu:35)
        self.L = lib(patched64)
        self.patched64 = patched64
        self.center = tuple(float(c) for c in center)
        self.half_edge = float(np.float32(half_edge))
        self.max_depth = int(max_depth)
        self._h = self.L.ref_create(_f(self.center), self.half_edge, self.max_depth)

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ref_destroy(self._h)
            self._h = None

    @property
    def size(self):
        return self.L.ref_size(self._h)

    def pool(self):
        out = np.zeros(2 * self.size, dtype=np.uint32)
        if self.size:
            self.L.ref_download_pool(self._h, _p(out))
        return out

    def load(self, pool):
        pool = np.ascontiguousarray(pool, dtype=np.uint32)
        self.L.ref_upload_pool(self._h, _p(pool), pool.size // 2)

    def integrate_depth(self, depth, rgb, fx, fy, pose=IDENTITY):
        """main.cpp:38-44 from host buffers; returns (ms of the reference path, ms incl. H2D)."""
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w = depth.shape
        ms, ms2 = C.c_double(), C.c_double()
        rc = self.L.ref_integrate_depth(self._h, _p(depth), _p(rgb), w, h, fx, fy, _f(mat_colmajor(pose)),
                                        C.byref(ms), C.byref(ms2))
        assert rc == 0, "reference CUDA error %d" % rc
        return ms.value, ms2.value

    def integrate_depth_dev(self, d_depth_ptr, d_rgb_ptr, w, h, fx, fy, pose=IDENTITY):
        ms = C.c_double()
        rc = self.L.ref_integrate_depth_dev(self._h, d_depth_ptr, d_rgb_ptr, w, h, fx, fy, _f(mat_colmajor(pose)),
                                            C.byref(ms))
        assert rc == 0, "reference CUDA error %d" % rc
        return ms.value

    def integrate_points(self, xyz, rgb):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        rc = self.L.ref_integrate_points(self._h, _p(xyz), _p(rgb), xyz.shape[0])
        assert rc == 0, "reference CUDA error %d" % rc

    def integrate_voxels(self, centers4, colors4):
        c = np.ascontiguousarray(centers4, dtype=np.float32)
        k = np.ascontiguousarray(colors4, dtype=np.float32)
        rc = self.L.ref_integrate_voxels(self._h, _p(c), _p(k), c.shape[0])
        assert rc == 0, "reference CUDA error %d" % rc

    def extract_voxels(self, max_depth=None):
        D = self.max_depth if max_depth is None else max_depth
        cap = max(self.size, 1)
        centers = np.zeros((cap, 4), dtype=np.float32)
        colors = np.zeros((cap, 4), dtype=np.float32)
        n = self.L.ref_extract_voxels(self._h, D, _p(centers), _p(colors), cap)
        return centers[:n], colors[:n]

    def raycast(self, w, h, fov=45.0, view=IDENTITY, want_image=True):
        out = np.zeros((h, w, 4), dtype=np.uint8)
        ms = C.c_double()
        rc = self.L.ref_raycast(self._h, _p(out) if want_image else None, w, h, float(fov),
                                _f(mat_colmajor(view)), C.byref(ms))
        assert rc == 0, "reference CUDA error %d" % rc
        return out, ms.value


def vertex_map(depth, fx, fy, pose=None, patched64=False):
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = depth.shape
    out = np.zeros((h * w, 3), dtype=np.float32)
    rc = lib(patched64).ref_vertex_map(_p(depth), w, h, fx, fy,
                                       _f(mat_colmajor(pose)) if pose is not None else None, _p(out))
    assert rc == 0
    return out


def bbox(points, init=None, patched64=False):
    p = np.ascontiguousarray(points, dtype=np.float32)
    b = _f(init if init is not None else [0] * 6)
    rc = lib(patched64).ref_bbox(_p(p), p.shape[0], b)
    assert rc == 0
    return np.array(list(b), dtype=np.float32)


# ---- camera tracking: the reference's image / localization kernels and its RGBDCamera -----------------------------
def bilateral(depth):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    out = np.empty_like(d)
    assert lib().ref_bilateral(_p(d), d.shape[1], d.shape[0], _p(out)) == 0
    return out


def subsample_depth(depth):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = d.shape
    out = np.empty((h // 2, w // 2), dtype=np.uint16)
    assert lib().ref_subsample_depth(_p(d), w, h, _p(out)) == 0
    return out


def subsample_f32(img):
    a = np.ascontiguousarray(img, dtype=np.float32)
    h, w = a.shape
    out = np.empty((h // 2, w // 2), dtype=np.float32)
    assert lib().ref_subsample_f32(_p(a), w, h, _p(out)) == 0
    return out


def normal_map(vtx, w, h):
    v = np.ascontiguousarray(vtx, dtype=np.float32).reshape(h * w, 3)
    out = np.empty_like(v)
    assert lib().ref_normal_map(_p(v), w, h, _p(out)) == 0
    return out


def transform_normals(nrm, M):
    p = np.ascontiguousarray(nrm, dtype=np.float32).copy()
    assert lib().ref_transform_normals(_p(p), p.shape[0], _f(mat_colmajor(M))) == 0
    return p


def color_to_intensity(rgb):
    c = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
    out = np.empty(c.shape[0], dtype=np.float32)
    assert lib().ref_color_to_intensity(_p(c), c.shape[0], _p(out)) == 0
    return out


def icp_cost2(last_v, last_n, cur_v, cur_n, w, h):
    arrs = [np.ascontiguousarray(a, dtype=np.float32).reshape(h * w, 3) for a in (last_v, last_n, cur_v, cur_n)]
    A, b = (C.c_float * 36)(), (C.c_float * 6)()
    assert lib().ref_icp_cost2(*[_p(a) for a in arrs], w, h, A, b) == 0
    return np.array(A, dtype=np.float32).reshape(6, 6), np.array(b, dtype=np.float32)


class RefTracker:
    """The reference's own sensor::RGBDCamera (rgbd_camera.cpp), fed with host depth frames."""

    def __init__(self, w, h, fx, fy):
        self._h = lib().ref_tracker_create(int(w), int(h), float(fx), float(fy))
        self.w, self.h = int(w), int(h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_tracker_destroy(self._h)
            self._h = None

    def update(self, depth):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        assert d.shape == (self.h, self.w)
        return lib().ref_tracker_update(self._h, _p(d))  # milliseconds

    def position(self):
        p, o = (C.c_float * 3)(), (C.c_float * 9)()
        lib().ref_tracker_pose(self._h, p, o)
        return np.array(p, dtype=np.float32)

    def orientation(self):
        p, o = (C.c_float * 3)(), (C.c_float * 9)()
        lib().ref_tracker_pose(self._h, p, o)
        return np.array(o, dtype=np.float32).reshape(3, 3).T
