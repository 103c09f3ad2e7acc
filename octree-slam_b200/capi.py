"""ctypes binding of libosl_b200.so (include/osl_b200.h) -- the same stub a reference maintainer would write.

No torch types cross the boundary: device buffers are passed as raw pointers (torch tensors' data_ptr(), or any
CUDA allocation), host buffers as numpy arrays.  If the CUDA library is missing this module raises: there is no CPU
fallback in the product path.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libosl_b200.so")
MAX_DEPTH = 20

OSL_OK = 0
STATUS = {0: "OSL_OK", -1: "OSL_ERR_INVALID", -2: "OSL_ERR_CUDA", -3: "OSL_ERR_OOM",
          -4: "OSL_ERR_POOL_OVERFLOW", -5: "OSL_ERR_UNSUPPORTED"}

EXPORTS = [
    "osl_svo_create", "osl_svo_destroy", "osl_svo_reset", "osl_svo_expand", "osl_svo_max_depth", "osl_svo_set_quirks", "osl_svo_set_pipeline", "osl_svo_set_stage_timing", "osl_get_stage_times",
    "osl_integrate_depth", "osl_integrate_depth_posed", "osl_integrate_depth_host", "osl_integrate_points", "osl_integrate_voxels",
    "osl_svo_sync", "osl_svo_join", "osl_svo_view", "osl_svo_size", "osl_svo_download", "osl_svo_upload", "osl_get_counters", "osl_svo_save", "osl_svo_load",
    "osl_raycast", "osl_raycast_host", "osl_raycast_pool", "osl_raycast_rows", "osl_raycast_bands", "osl_extract_voxels", "osl_voxelize_mesh", "osl_voxelize_thin", "osl_free_device", "osl_copy_device", "osl_debug_trace",
    "osl_svo_reserve", "osl_svo_pool_device", "osl_svo_adopt", "osl_svo_delta_bytes", "osl_svo_delta_pack", "osl_svo_delta_apply",
    "osl_shard_analyze", "osl_shard_assign", "osl_shard_delta_bytes", "osl_shard_delta_pack", "osl_shard_delta_apply", "osl_shard_fixup",
    "osl_generate_vertex_map", "osl_transform_vertex_map", "osl_point_cloud_bbox", "osl_compute_keys",
    "osl_bilateral_filter", "osl_subsample_depth", "osl_subsample_f32", "osl_generate_normal_map", "osl_transform_normal_map",
    "osl_color_to_intensity", "osl_icp_cost",
    "osl_tracker_create", "osl_tracker_destroy", "osl_tracker_reset", "osl_tracker_update", "osl_tracker_update_host",
    "osl_tracker_get_pose", "osl_tracker_pose_device", "osl_tracker_view",
    "osl_status_string", "osl_last_cuda_error", "osl_version", "osl_frame_result_bytes", "osl_launch_count", "osl_debug_profile", "osl_debug_cta_profile", "osl_debug_scramble_hints",
]


class OslError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        super().__init__("%s failed: %s (cuda error %d)" % (where, STATUS.get(status, status),
                                                            lib().osl_last_cuda_error()))


class Counters(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64), ("n_valid", C.c_int64), ("n_unique", C.c_int64), ("n_split", C.c_int64),
        ("pass_sizes", C.c_int64 * (MAX_DEPTH + 1)), ("parents", C.c_int64 * (MAX_DEPTH + 1)),
        ("n_nodes", C.c_int64), ("algorithmic_bytes", C.c_int64), ("total_algorithmic_bytes", C.c_int64),
        ("frames", C.c_int64),
    ]


class RaycastParams(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("start_dist", C.c_float), ("max_range", C.c_float),
                ("mode", C.c_int)]


class RaycastStats(C.Structure):
    _fields_ = [("rays", C.c_int64), ("steps", C.c_int64), ("visits", C.c_int64)]


_lib = None


def lib():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libosl_b200.so is not built: run `python __graft_entry__.py build` "
                          "(the product path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    fp = C.POINTER(C.c_float)
    sig = {
        "osl_svo_create": (i32, [C.POINTER(vp), fp, f32, i32, C.c_size_t, i32]),
        "osl_svo_destroy": (None, [vp]),
        "osl_svo_reset": (i32, [vp]),
        "osl_svo_expand": (i32, [vp, i32]),
        "osl_svo_max_depth": (i32, [vp]),
        "osl_svo_set_quirks": (i32, [vp, i32]),
        "osl_svo_set_pipeline": (i32, [vp, i32]),
        "osl_svo_set_stage_timing": (i32, [vp, i32]),
        "osl_get_stage_times": (i32, [vp, fp]),
        "osl_integrate_depth": (i32, [vp, vp, vp, i32, i32, f32, f32, fp, vp]),
        "osl_integrate_depth_posed": (i32, [vp, vp, vp, i32, i32, f32, f32, vp, vp]),
        "osl_integrate_depth_host": (i32, [vp, vp, vp, i32, i32, f32, f32, fp, vp]),
        "osl_integrate_points": (i32, [vp, vp, vp, i32, vp]),
        "osl_integrate_voxels": (i32, [vp, vp, vp, i32, vp]),
        "osl_svo_sync": (i32, [vp]),
        "osl_svo_join": (i32, [vp, vp]),
        "osl_svo_view": (i32, [vp, C.POINTER(vp), C.POINTER(i32), fp, fp]),
        "osl_svo_size": (i32, [vp]),
        "osl_svo_download": (i32, [vp, vp, i32]),
        "osl_svo_upload": (i32, [vp, vp, i32]),
        "osl_get_counters": (i32, [vp, C.POINTER(Counters)]),
        "osl_svo_save": (i32, [vp, C.c_char_p]),
        "osl_svo_load": (i32, [vp, C.c_char_p]),
        "osl_raycast": (i32, [vp, vp, i32, i32, f32, fp, C.POINTER(RaycastParams), vp]),
        "osl_raycast_host": (i32, [vp, vp, i32, i32, f32, fp, C.POINTER(RaycastParams), C.POINTER(RaycastStats), vp]),
        "osl_raycast_rows": (i32, [vp, vp, i32, i32, i32, i32, f32, fp, C.POINTER(RaycastParams),
                                   C.POINTER(RaycastStats), vp]),
        "osl_raycast_bands": (i32, [vp, vp, i32, i32, i32, i32, i32, f32, fp, C.POINTER(RaycastParams),
                                    C.POINTER(i32), vp]),
        "osl_raycast_pool": (i32, [vp, fp, f32, vp, i32, i32, f32, fp, C.POINTER(RaycastParams),
                                   C.POINTER(RaycastStats), vp]),
        "osl_extract_voxels": (i32, [vp, i32, vp, vp, vp, i64, C.POINTER(i64), vp]),
        "osl_voxelize_mesh": (i32, [vp, i32, vp, i32, vp, fp, f32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                    C.POINTER(vp), C.POINTER(i64), vp]),
        "osl_voxelize_thin": (i32, [vp, i32, vp, i32, vp, fp, fp, i32, fp, f32, i32, C.POINTER(vp), C.POINTER(vp),
                                    C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), vp]),
        "osl_free_device": (None, [vp]),
        "osl_copy_device": (i32, [vp, vp, C.c_size_t]),
        "osl_debug_trace": (i32, [vp, i32, vp]),
        "osl_svo_reserve": (i32, [vp, C.c_size_t]),
        "osl_svo_pool_device": (i32, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "osl_svo_adopt": (i32, [vp, i32, i32, C.POINTER(C.c_float), C.c_float]),
        "osl_svo_delta_bytes": (C.c_size_t, [vp]),
        "osl_svo_delta_pack": (i32, [vp, vp, C.c_size_t, C.POINTER(C.c_size_t), vp]),
        "osl_svo_delta_apply": (i32, [vp, vp, C.c_size_t, vp]),
        "osl_shard_analyze": (i32, [vp, vp, i32, i32, i32, vp, C.POINTER(i32), vp]),
        "osl_shard_assign": (i32, [vp, vp, vp, vp, vp]),
        "osl_shard_delta_bytes": (C.c_size_t, [vp]),
        "osl_shard_delta_pack": (i32, [vp, vp, C.c_size_t, C.POINTER(C.c_size_t), vp]),
        "osl_shard_delta_apply": (i32, [vp, vp, C.c_size_t, vp]),
        "osl_shard_fixup": (i32, [vp, vp, i32, vp, i32, vp]),
        "osl_generate_vertex_map": (i32, [vp, vp, i32, i32, f32, f32, i32, i32, vp]),
        "osl_transform_vertex_map": (i32, [vp, fp, i32, vp]),
        "osl_point_cloud_bbox": (i32, [vp, i32, fp, vp]),
        "osl_compute_keys": (i32, [vp, i32, i32, fp, f32, i32, vp, vp]),
        "osl_bilateral_filter": (i32, [vp, vp, i32, i32, vp]),
        "osl_subsample_depth": (i32, [vp, vp, i32, i32, vp]),
        "osl_subsample_f32": (i32, [vp, vp, i32, i32, vp]),
        "osl_generate_normal_map": (i32, [vp, vp, i32, i32, vp]),
        "osl_transform_normal_map": (i32, [vp, fp, i32, vp]),
        "osl_color_to_intensity": (i32, [vp, vp, i32, vp]),
        "osl_icp_cost": (i32, [vp, vp, vp, vp, i32, i32, fp, fp, C.POINTER(i32), vp]),
        "osl_tracker_create": (i32, [C.POINTER(vp), i32, i32, f32, f32, i32, i32]),
        "osl_tracker_destroy": (None, [vp]),
        "osl_tracker_reset": (i32, [vp]),
        "osl_tracker_update": (i32, [vp, vp, vp]),
        "osl_tracker_update_host": (i32, [vp, vp, vp]),
        "osl_tracker_get_pose": (i32, [vp, fp, fp, fp, C.POINTER(i32), C.POINTER(i32)]),
        "osl_tracker_pose_device": (i32, [vp, C.POINTER(vp)]),
        "osl_tracker_view": (i32, [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]),
        "osl_status_string": (C.c_char_p, [i32]),
        "osl_last_cuda_error": (i32, []),
        "osl_version": (C.c_char_p, []),
        "osl_frame_result_bytes": (i32, []),
        "osl_launch_count": (i64, []),
        "osl_debug_profile": (i32, [vp, i32]),
        "osl_debug_cta_profile": (i32, [vp]),
        "osl_debug_scramble_hints": (i32, [vp, C.c_uint]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return _lib


def _check(status, where):
    if status != OSL_OK:
        raise OslError(status, where)


def _f(arr):
    arr = [float(x) for x in arr]
    return (C.c_float * len(arr))(*arr)


def mat_colmajor(m):
    """4x4 matrix in math convention (m[r][c]) -> 16 floats, column-major (glm::mat4 memory order)."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T).reshape(16)


IDENTITY = np.eye(4, dtype=np.float32)


def _hptr(a):
    return a.ctypes.data_as(C.c_void_p)
