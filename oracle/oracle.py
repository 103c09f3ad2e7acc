"""ctypes binding of the CPU oracle (oracle/osl_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libosl_oracle.so")
MAX_DEPTH = 20


SOURCES = ["osl_oracle.c", "osl_oracle_track.c", "osl_oracle_thin.c"]


def build(force=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC"] + srcs +
        ["-o", LIB, "-lm"])
    return LIB


class Counters(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64), ("n_valid", C.c_int64), ("n_unique", C.c_int64),
        ("n_split", C.c_int64),
        ("pass_sizes", C.c_int64 * (MAX_DEPTH + 1)),
        ("parents", C.c_int64 * (MAX_DEPTH + 1)),
        ("ray_steps", C.c_int64), ("ray_visits", C.c_int64),
    ]


class RaycastParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("start_dist", C.c_float),
                ("max_range", C.c_float), ("mode", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        f3 = C.POINTER(C.c_float)
        L.orc_compute_key.restype = C.c_int64
        L.orc_compute_key.argtypes = [C.c_float, C.c_float, C.c_float, f3, C.c_int, C.c_float]
        L.orc_compute_keys.argtypes = [C.c_void_p, C.c_int, C.c_int, f3, C.c_float, C.c_int, C.c_void_p]
        L.orc_vertex_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_transform.argtypes = [C.c_void_p, C.c_int, f3]
        L.orc_bbox.argtypes = [C.c_void_p, C.c_int, f3]
        L.orc_svo_create.restype = C.c_void_p
        L.orc_svo_create.argtypes = [f3, C.c_float, C.c_int, C.c_int]
        L.orc_svo_destroy.argtypes = [C.c_void_p]
        L.orc_svo_size.argtypes = [C.c_void_p]
        L.orc_svo_pool.restype = C.c_void_p
        L.orc_svo_pool.argtypes = [C.c_void_p]
        L.orc_svo_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.orc_svo_expand.argtypes = [C.c_void_p, C.c_int]
        L.orc_svo_half_edge.restype = C.c_float
        L.orc_svo_half_edge.argtypes = [C.c_void_p]
        L.orc_svo_max_depth.argtypes = [C.c_void_p]
        L.orc_svo_load.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_integrate_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_integrate_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_float, C.c_float, f3]
        L.orc_integrate_voxels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_extract_voxels.restype = C.c_int64
        L.orc_extract_voxels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_raycast.argtypes = [C.c_void_p, f3, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                  f3, C.POINTER(RaycastParams), C.POINTER(Counters), C.c_int64]
        L.orc_voxelize_mesh.restype = C.c_int64
        L.orc_voxelize_mesh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, f3, C.c_float, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        u16p = C.c_void_p
        L.orc_bilateral.argtypes = [u16p, u16p, C.c_int, C.c_int]
        L.orc_subsample_depth.argtypes = [u16p, u16p, C.c_int, C.c_int]
        L.orc_normal_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_transform_normals.argtypes = [C.c_void_p, C.c_int, f3]
        L.orc_color_to_intensity.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_subsample_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_icp_cost.restype = C.c_int64
        L.orc_icp_cost.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, f3, f3]
        L.orc_solve_cholesky.argtypes = [C.c_int, f3, f3, f3]
        L.orc_pose_increment.argtypes = [f3, C.c_int, f3]
        L.orc_tracker_create.restype = C.c_void_p
        L.orc_tracker_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_tracker_destroy.argtypes = [C.c_void_p]
        L.orc_tracker_update.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_tracker_pose.argtypes = [C.c_void_p, f3, f3, f3]
        L.orc_tracker_lost.argtypes = [C.c_void_p]
        L.orc_tracker_pairs.restype = C.c_int64
        L.orc_tracker_pairs.argtypes = [C.c_void_p]
        L.orc_nv_logf.restype = C.c_float
        L.orc_nv_logf.argtypes = [C.c_float]
        L.orc_mat4_inverse.argtypes = [f3, f3]
        _lib = L
    return _lib


def _f(arr):
    return (C.c_float * len(arr))(*[float(x) for x in arr])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


IDENTITY = np.eye(4, dtype=np.float32)


def mat_colmajor(m):
    """4x4 numpy matrix (math convention, m[r][c]) -> 16 floats column-major (glm)."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T).reshape(16)


def compute_keys(points, center, half_edge, max_depth):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, stride = pts.shape
    keys = np.empty(n, dtype=np.int64)
    lib().orc_compute_keys(_ptr(pts), stride, n, _f(center), half_edge, max_depth, _ptr(keys))
    return keys


def vertex_map(depth, fx, fy):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = d.shape
    out = np.empty((h * w, 3), dtype=np.float32)
    lib().orc_vertex_map(_ptr(d), _ptr(out), w, h, fx, fy, w, h)
    return out


def transform(points, pose):
    p = np.ascontiguousarray(points, dtype=np.float32).copy()
    lib().orc_transform(_ptr(p), p.shape[0], _f(mat_colmajor(pose)))
    return p


def bbox(points, init=None):
    p = np.ascontiguousarray(points, dtype=np.float32)
    b = _f(init if init is not None else [0] * 6)
    lib().orc_bbox(_ptr(p), p.shape[0], b)
    return np.array(list(b), dtype=np.float32)


class OracleSVO:
    """CPU restatement of the reference's Octree + svo.cu pool (one GPU sub-tree)."""

    def __init__(self, center=(0, 0, 0), half_edge=1.0, max_depth=8, quirks=True):
        self.center = tuple(float(c) for c in center)
        self.half_edge = float(np.float32(half_edge))
        self.max_depth = int(max_depth)
        self._h = lib().orc_svo_create(_f(self.center), self.half_edge, self.max_depth, int(quirks))
        if not self._h:
            raise ValueError("bad max_depth")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_svo_destroy(self._h)
            self._h = None

    @property
    def size(self):
        return lib().orc_svo_size(self._h)

    def pool(self):
        n = self.size
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        p = lib().orc_svo_pool(self._h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(2 * n,)).copy()

    def load(self, pool):
        pool = np.ascontiguousarray(pool, dtype=np.uint32)
        lib().orc_svo_load(self._h, _ptr(pool), pool.size // 2)

    def expand(self, layers=1):
        if lib().orc_svo_expand(self._h, int(layers)) != 0:
            raise ValueError("bad layers")
        self.half_edge = float(lib().orc_svo_half_edge(self._h))
        self.max_depth = int(lib().orc_svo_max_depth(self._h))

    def counters(self):
        c = Counters()
        lib().orc_svo_counters(self._h, C.byref(c))
        return c

    def integrate_points(self, xyz, rgb):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        assert xyz.shape[1] == 3 and rgb.shape == (xyz.shape[0], 3)
        lib().orc_integrate_points(self._h, _ptr(xyz), _ptr(rgb), xyz.shape[0])

    def integrate_depth(self, depth, rgb, fx, fy, pose=IDENTITY):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = d.shape
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(h * w, 3)
        lib().orc_integrate_depth(self._h, _ptr(d), _ptr(rgb), w, h, fx, fy, _f(mat_colmajor(pose)))

    def integrate_voxels(self, centers4, colors4):
        c = np.ascontiguousarray(centers4, dtype=np.float32)
        k = np.ascontiguousarray(colors4, dtype=np.float32)
        assert c.shape[1] == 4 and k.shape == c.shape
        lib().orc_integrate_voxels(self._h, _ptr(c), _ptr(k), c.shape[0])

    def extract_voxels(self, max_depth=None):
        D = self.max_depth if max_depth is None else max_depth
        n = lib().orc_extract_voxels(self._h, D, None, None, None, 0)
        centers = np.empty((n, 4), dtype=np.float32)
        colors = np.empty((n, 4), dtype=np.float32)
        keys = np.empty(n, dtype=np.int64)
        lib().orc_extract_voxels(self._h, D, _ptr(centers), _ptr(colors), _ptr(keys), n)
        return centers, colors, keys

    def raycast(self, w, h, fov=45.0, view=IDENTITY, mode=0, max_steps=0, counters=None, params=None):
        return raycast(self.pool(), self.center, self.half_edge, w, h, fov, view, mode, max_steps,
                       counters, params)


def raycast(pool, center, half_edge, w, h, fov=45.0, view=IDENTITY, mode=0, max_steps=0,
            counters=None, params=None):
    pool = np.ascontiguousarray(pool, dtype=np.uint32)
    out = np.zeros((h, w, 4), dtype=np.uint8)
    prm = params if params is not None else RaycastParams(532.57, 531.54, 0.002, 10.0, mode)
    prm.mode = mode
    lib().orc_raycast(_ptr(pool), _f(center), float(half_edge), _ptr(out), w, h, float(fov),
                      _f(mat_colmajor(view)), C.byref(prm),
                      C.byref(counters) if counters is not None else None, int(max_steps))
    return out


def voxelize_mesh(vertices, triangles, center, half_edge, max_depth):
    """-> (keys int64[n] ascending leading-1 Morton keys, tris int32[n] lowest triangle per cell, centers float32[n,4])"""
    V = np.ascontiguousarray(vertices, dtype=np.float32)
    T = np.ascontiguousarray(triangles, dtype=np.int32)
    n = lib().orc_voxelize_mesh(_ptr(V), V.shape[0], _ptr(T), T.shape[0], _f(center), float(half_edge),
                                int(max_depth), None, None, None, 0)
    keys = np.empty(n, dtype=np.int64)
    tris = np.empty(n, dtype=np.int32)
    cen = np.empty((n, 4), dtype=np.float32)
    lib().orc_voxelize_mesh(_ptr(V), V.shape[0], _ptr(T), T.shape[0], _f(center), float(half_edge),
                            int(max_depth), _ptr(keys), _ptr(tris), _ptr(cen), n)
    return keys, tris, cen


def voxelize_thin(vertices, triangles, bbox0, bbox1, log_n=8):
    """The reference's voxelisation rule (voxelpipe THIN_RASTER on the dense 2^log_n grid over the mesh bounding box,
    voxelization.cu:24,281-285) -> (cells int32[n,3] (x, y, z) in ascending z,y,x order, tris int32[n] lowest triangle)"""
    V = np.ascontiguousarray(vertices, dtype=np.float32)
    T = np.ascontiguousarray(triangles, dtype=np.int32)
    L = lib()
    L.orc_voxelize_thin.restype = C.c_longlong
    args = [_ptr(V), C.c_int(V.shape[0]), _ptr(T), C.c_int(T.shape[0]), _f(bbox0), _f(bbox1), C.c_int(int(log_n))]
    n = L.orc_voxelize_thin(*args, None, None, C.c_longlong(0))
    cells = np.empty((n, 3), dtype=np.int32)
    tris = np.empty(n, dtype=np.int32)
    L.orc_voxelize_thin(*args, _ptr(cells), _ptr(tris), C.c_longlong(n))
    return cells, tris


# ---------------------------------------------------------------------------------------------- camera tracking
def bilateral(depth):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    out = np.empty_like(d)
    lib().orc_bilateral(_ptr(d), _ptr(out), d.shape[1], d.shape[0])
    return out


def subsample_depth(depth):
    d = np.ascontiguousarray(depth, dtype=np.uint16)
    h, w = d.shape
    out = np.empty((h // 2, w // 2), dtype=np.uint16)
    lib().orc_subsample_depth(_ptr(d), _ptr(out), w, h)
    return out


def normal_map(vtx, w, h):
    v = np.ascontiguousarray(vtx, dtype=np.float32).reshape(h * w, 3)
    out = np.empty_like(v)
    lib().orc_normal_map(_ptr(v), _ptr(out), w, h)
    return out


def transform_normals(nrm, M):
    p = np.ascontiguousarray(nrm, dtype=np.float32).copy()
    lib().orc_transform_normals(_ptr(p), p.shape[0], _f(mat_colmajor(M)))
    return p


def color_to_intensity(rgb):
    c = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
    out = np.empty(c.shape[0], dtype=np.float32)
    lib().orc_color_to_intensity(_ptr(c), _ptr(out), c.shape[0])
    return out


def subsample_f32(img):
    a = np.ascontiguousarray(img, dtype=np.float32)
    h, w = a.shape
    out = np.empty((h // 2, w // 2), dtype=np.float32)
    lib().orc_subsample_f32(_ptr(a), _ptr(out), w, h)
    return out


def icp_cost(last_v, last_n, cur_v, cur_n, exact_jacobian=False):
    """-> (A 6x6, b 6, pairs): localization_kernels.cu computeICPCost2, sums in double"""
    arrs = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (last_v, last_n, cur_v, cur_n)]
    A, b = (C.c_float * 36)(), (C.c_float * 6)()
    pairs = lib().orc_icp_cost(*[_ptr(a) for a in arrs], arrs[0].shape[0], int(exact_jacobian), A, b)
    return np.array(A, dtype=np.float32).reshape(6, 6), np.array(b, dtype=np.float32), int(pairs)


def solve_cholesky(A, b):
    x = (C.c_float * 6)()
    lib().orc_solve_cholesky(6, _f(np.asarray(A, dtype=np.float32).reshape(36)), _f(b), x)
    return np.array(x, dtype=np.float32)


def pose_increment(x, exact_jacobian=False):
    out = (C.c_float * 16)()
    lib().orc_pose_increment(_f(x), int(exact_jacobian), out)
    return np.array(out, dtype=np.float32).reshape(4, 4).T  # math convention m[r][c]


class OracleTracker:
    """CPU restatement of sensor::RGBDCamera (rgbd_camera.cpp:53-191)."""

    def __init__(self, w, h, fx, fy, exact_jacobian=False):
        self._h = lib().orc_tracker_create(int(w), int(h), float(fx), float(fy), int(exact_jacobian))
        self.w, self.h = int(w), int(h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tracker_destroy(self._h)
            self._h = None

    def update(self, depth):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        assert d.shape == (self.h, self.w)
        lib().orc_tracker_update(self._h, _ptr(d))

    def pose(self):
        """main.cpp:40's mat4(orientation) * translate(position), math convention m[r][c]"""
        m = (C.c_float * 16)()
        lib().orc_tracker_pose(self._h, m, None, None)
        return np.array(m, dtype=np.float32).reshape(4, 4).T

    def position(self):
        p = (C.c_float * 3)()
        lib().orc_tracker_pose(self._h, None, p, None)
        return np.array(p, dtype=np.float32)

    def orientation(self):
        o = (C.c_float * 9)()
        lib().orc_tracker_pose(self._h, None, None, o)
        return np.array(o, dtype=np.float32).reshape(3, 3).T

    @property
    def lost(self):
        return bool(lib().orc_tracker_lost(self._h))

    @property
    def pairs(self):
        return int(lib().orc_tracker_pairs(self._h))
