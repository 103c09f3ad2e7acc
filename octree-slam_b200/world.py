"""Host-side Python mirror of the reference's world / rendering interface for the hot path, over the C ABI.

Names, argument meaning and behaviour follow the reference classes so tests read like the reference's call sites:
  world::Octree  (include/octree_slam/world/octree.h:83-126, src/world/octree.cpp:251-389)
  world::Scene   (include/octree_slam/world/scene.h:22-78,  src/world/scene.cpp:98-113)
  rendering::coneTraceSVO (include/octree_slam/rendering/cone_tracing_kernels.h:16)
torch is used only to hold device buffers; every computation happens in libosl_b200.so.
"""
import collections
import ctypes as C
import math

import numpy as np

from . import capi
from .capi import IDENTITY, Counters, RaycastParams, RaycastStats, _check, _f, _hptr, lib, mat_colmajor


OSL_NCOUNT_MAX = 20 + 21 * 21  # OSL_NCOUNT(OSL_MAX_DEPTH): level counters + (frontier, depth) bucket counters


def _torch():
    import torch
    return torch


def _dev(arr, dtype, device):
    """numpy / torch -> contiguous CUDA torch tensor of dtype."""
    torch = _torch()
    if isinstance(arr, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype))
        return t.to("cuda:%d" % device, non_blocking=False)
    assert arr.is_cuda and arr.is_contiguous()
    return arr


class SVO:
    """One GPU-resident sparse voxel octree (the reference's OctreeNode::gpu_data_ / gpu_size_)."""

    def __init__(self, center=(0.0, 0.0, 0.0), half_edge=1.0, max_depth=8, reserve_nodes=0, device=0,
                 quirks=True, force_grid_sort=False, zero_copy=False, fused=True):
        self.center = tuple(float(c) for c in center)
        self.half_edge = float(np.float32(half_edge))
        self.max_depth = int(max_depth)
        self.device = device
        h = C.c_void_p()
        _check(lib().osl_svo_create(C.byref(h), _f(self.center), self.half_edge, self.max_depth,
                                    int(reserve_nodes), device), "osl_svo_create")
        self._h = h
        # integrate calls are asynchronous and up to 8 frames (the library's result ring, OSL_RING) are in flight on the
        # library's own streams: the input buffers of the last 2 * 8 calls stay referenced, so that torch's caching
        # allocator cannot hand a block a queued kernel still reads to a later copy
        self._keep = collections.deque(maxlen=16)
        if not quirks or force_grid_sort or zero_copy or not fused:
            _check(lib().osl_svo_set_quirks(self._h, int(bool(quirks)) | (2 if force_grid_sort else 0) |
                                            (4 if zero_copy else 0) | (0 if fused else 8)),
                   "osl_svo_set_quirks")

    def set_pipeline(self, enabled=True):
        """inputs of later integrate calls are complete at call time -> frame f+1's sort overlaps frame f's tree update"""
        _check(lib().osl_svo_set_pipeline(self._h, int(bool(enabled))), "osl_svo_set_pipeline")
        return self

    def set_stage_timing(self, enabled=True):
        _check(lib().osl_svo_set_stage_timing(self._h, int(bool(enabled))), "osl_svo_set_stage_timing")
        return self

    def stage_times(self):
        """ms of (k_emit, k_sort, k_structure, k_levels) of the last timed (non-pipelined) frame"""
        ms = (C.c_float * 4)()
        _check(lib().osl_get_stage_times(self._h, ms), "osl_get_stage_times")
        return [float(x) for x in ms]

    def close(self):
        if getattr(self, "_h", None):
            lib().osl_svo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- integrate -------------------------------------------------------------------------------------
    def integrate_depth(self, depth, rgb, fx, fy, pose=IDENTITY, stream=None):
        """depth: (H,W) uint16 mm, rgb: (H,W,3) uint8; numpy (copied to the device) or CUDA torch tensors."""
        h, w = depth.shape[:2]
        d = _dev(depth, np.uint16, self.device)
        c = _dev(rgb, np.uint8, self.device)
        _check(lib().osl_integrate_depth(self._h, d.data_ptr(), c.data_ptr(), w, h, fx, fy,
                                         _f(mat_colmajor(pose)), stream), "osl_integrate_depth")
        self._keep.append((d, c))
        return self

    def integrate_depth_tracked(self, depth, rgb, fx, fy, camera, stream=None):
        """main.cpp:33-44 with the tracking line live and no host round trip: `camera` (sensor.RGBDCamera) localises
        the frame on `stream`, its pose stays in device memory and drives the integration of the same frame."""
        h, w = depth.shape[:2]
        d = _dev(depth, np.uint16, self.device)
        c = _dev(rgb, np.uint8, self.device)
        _check(lib().osl_tracker_update(camera._h, d.data_ptr(), stream), "osl_tracker_update")
        _check(lib().osl_integrate_depth_posed(self._h, d.data_ptr(), c.data_ptr(), w, h, fx, fy,
                                               camera.pose_device(), stream), "osl_integrate_depth_posed")
        self._keep.append((d, c))
        return self

    def integrate_depth_host(self, depth, rgb, fx, fy, pose=IDENTITY, stream=None):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w = depth.shape
        _check(lib().osl_integrate_depth_host(self._h, _hptr(depth), _hptr(rgb), w, h, fx, fy,
                                              _f(mat_colmajor(pose)), stream), "osl_integrate_depth_host")
        self._keep.append((depth, rgb))
        return self

    def integrate_points(self, xyz, rgb, stream=None):
        p = _dev(xyz, np.float32, self.device)
        c = _dev(rgb, np.uint8, self.device)
        n = p.shape[0]
        _check(lib().osl_integrate_points(self._h, p.data_ptr() if n else None, c.data_ptr() if n else None, n,
                                          stream), "osl_integrate_points")
        self._keep.append((p, c))
        return self

    def integrate_voxels(self, centers4, colors4, stream=None):
        p = _dev(centers4, np.float32, self.device)
        c = _dev(colors4, np.float32, self.device)
        n = p.shape[0]
        _check(lib().osl_integrate_voxels(self._h, p.data_ptr() if n else None, c.data_ptr() if n else None, n,
                                          stream), "osl_integrate_voxels")
        self._keep.append((p, c))
        return self

    # ---- views -----------------------------------------------------------------------------------------
    @property
    def size(self):
        return lib().osl_svo_size(self._h)

    def pool(self):
        n = self.size
        out = np.zeros(2 * n, dtype=np.uint32)
        if n:
            _check(lib().osl_svo_download(self._h, _hptr(out), n), "osl_svo_download")
        return out

    def load(self, pool):
        pool = np.ascontiguousarray(pool, dtype=np.uint32)
        _check(lib().osl_svo_upload(self._h, _hptr(pool), pool.size // 2), "osl_svo_upload")

    # ---- replicas (shard.replicate_tree / replicate_delta) ------------------------------------------------
    def reserve(self, n_nodes):
        _check(lib().osl_svo_reserve(self._h, int(n_nodes)), "osl_svo_reserve")

    def pool_device(self):
        """(device pointer of the pool, capacity in nodes); waits for the pipeline"""
        ptr, cap = C.c_void_p(), C.c_size_t()
        _check(lib().osl_svo_pool_device(self._h, C.byref(ptr), C.byref(cap)), "osl_svo_pool_device")
        return ptr.value, cap.value

    def adopt(self, n_nodes, max_depth, center, half_edge):
        """publish a pool a collective wrote into pool_device(): geometry must match, child pointers are validated"""
        _check(lib().osl_svo_adopt(self._h, int(n_nodes), int(max_depth), _f(center), float(half_edge)), "osl_svo_adopt")

    def delta_bytes(self):
        return int(lib().osl_svo_delta_bytes(self._h))

    def delta_pack(self, buf, cap_bytes, stream=None):
        n = C.c_size_t()
        _check(lib().osl_svo_delta_pack(self._h, buf.data_ptr(), int(cap_bytes), C.byref(n), stream), "osl_svo_delta_pack")
        return n.value

    def delta_apply(self, buf, nbytes, stream=None):
        _check(lib().osl_svo_delta_apply(self._h, buf.data_ptr(), int(nbytes), stream), "osl_svo_delta_apply")

    # ---- one map built by several GPUs (shard.integrate_voxels_sharded) -----------------------------------
    def shard_analyze(self, centers4, n_total, lo, hi, stream=None):
        tot = np.zeros(OSL_NCOUNT_MAX, dtype=np.uint32)
        nc = C.c_int()
        _check(lib().osl_shard_analyze(self._h, centers4.data_ptr(), int(n_total), int(lo), int(hi), _hptr(tot),
                                       C.byref(nc), stream), "osl_shard_analyze")
        return tot[:nc.value].copy()

    def shard_assign(self, colors4, base, totals, stream=None):
        base = np.ascontiguousarray(base, dtype=np.uint32)
        totals = np.ascontiguousarray(totals, dtype=np.uint32)
        _check(lib().osl_shard_assign(self._h, colors4.data_ptr(), _hptr(base), _hptr(totals), stream), "osl_shard_assign")

    def shard_delta_bytes(self):
        return int(lib().osl_shard_delta_bytes(self._h))

    def shard_delta_pack(self, buf, cap_bytes, stream=None):
        n = C.c_size_t()
        _check(lib().osl_shard_delta_pack(self._h, buf.data_ptr(), int(cap_bytes), C.byref(n), stream), "osl_shard_delta_pack")
        return n.value

    def shard_delta_apply(self, ptr, nbytes, stream=None):
        _check(lib().osl_shard_delta_apply(self._h, ptr, int(nbytes), stream), "osl_shard_delta_apply")

    def shard_fixup(self, centers4, n_total, starts, stream=None):
        st = np.ascontiguousarray(starts, dtype=np.int32)
        _check(lib().osl_shard_fixup(self._h, centers4.data_ptr(), int(n_total), _hptr(st) if st.size else None, int(st.size),
                                     stream), "osl_shard_fixup")

    def join(self, stream=None):
        """order `stream` after every frame enqueued so far (device-side)"""
        _check(lib().osl_svo_join(self._h, stream), "osl_svo_join")

    def save(self, path):
        _check(lib().osl_svo_save(self._h, str(path).encode()), "osl_svo_save")

    def restore(self, path):
        _check(lib().osl_svo_load(self._h, str(path).encode()), "osl_svo_load")

    def sync(self):
        _check(lib().osl_svo_sync(self._h), "osl_svo_sync")

    def reset(self):
        _check(lib().osl_svo_reset(self._h), "osl_svo_reset")

    def expand(self, layers=1):
        """Map growth (octree.cpp:183-206, 362-378; quirk Q10 fixed): double the half edge about the same centre
        `layers` times, one more level each, resolution kept."""
        _check(lib().osl_svo_expand(self._h, int(layers)), "osl_svo_expand")
        self.max_depth = lib().osl_svo_max_depth(self._h)
        self.half_edge = self.view()[3]

    def view(self):
        ptr, n = C.c_void_p(), C.c_int()
        c, he = (C.c_float * 3)(), C.c_float()
        _check(lib().osl_svo_view(self._h, C.byref(ptr), C.byref(n), c, C.byref(he)), "osl_svo_view")
        return ptr.value, n.value, tuple(c), he.value

    def counters(self):
        c = Counters()
        _check(lib().osl_get_counters(self._h, C.byref(c)), "osl_get_counters")
        return c

    # ---- raycast ---------------------------------------------------------------------------------------
    def raycast(self, w, h, fov=45.0, view=IDENTITY, mode=0, stats=None, params=None, stream=None):
        """Returns an (h, w, 4) uint8 numpy image {R,G,B,A} (host buffers through osl_raycast_host)."""
        out = np.zeros((h, w, 4), dtype=np.uint8)
        prm = params if params is not None else RaycastParams(532.57, 531.54, 0.002, 10.0, mode)
        prm.mode = mode
        _check(lib().osl_raycast_host(self._h, _hptr(out), w, h, float(fov), _f(mat_colmajor(view)), C.byref(prm),
                                      C.byref(stats) if stats is not None else None, stream), "osl_raycast_host")
        return out

    def raycast_device(self, out, w, h, fov=45.0, view=IDENTITY, mode=0, stream=None):
        prm = RaycastParams(532.57, 531.54, 0.002, 10.0, mode)
        _check(lib().osl_raycast(self._h, out.data_ptr(), w, h, float(fov), _f(mat_colmajor(view)), C.byref(prm),
                                 stream), "osl_raycast")
        return out

    def raycast_rows(self, out, w, h, row0, rows, fov=45.0, view=IDENTITY, mode=0, stream=None):
        """rows [row0, row0+rows) of the w x h image into the CUDA tensor `out` (rows x w x 4 uint8)"""
        prm = RaycastParams(532.57, 531.54, 0.002, 10.0, mode)
        _check(lib().osl_raycast_rows(self._h, out.data_ptr(), w, h, row0, rows, float(fov), _f(mat_colmajor(view)),
                                      C.byref(prm), None, stream), "osl_raycast_rows")
        return out

    def raycast_bands(self, out, w, h, band_h, n_ranks, rank, fov=45.0, view=IDENTITY, mode=0, stream=None):
        """every row `rank` owns (interleaved bands of band_h rows over n_ranks ranks) in one launch, compactly into
        the CUDA tensor `out`; returns the number of rows rendered"""
        prm = RaycastParams(532.57, 531.54, 0.002, 10.0, mode)
        rows = C.c_int()
        _check(lib().osl_raycast_bands(self._h, out.data_ptr(), w, h, band_h, n_ranks, rank, float(fov),
                                       _f(mat_colmajor(view)), C.byref(prm), C.byref(rows), stream),
               "osl_raycast_bands")
        return rows.value

    # ---- extraction ------------------------------------------------------------------------------------
    def extract_voxels(self, max_depth=None):
        torch = _torch()
        D = self.max_depth if max_depth is None else int(max_depth)
        n = C.c_int64()
        _check(lib().osl_extract_voxels(self._h, D, None, None, None, 0, C.byref(n), None), "osl_extract_voxels")
        cnt = n.value
        dev = "cuda:%d" % self.device
        centers = torch.empty((max(cnt, 1), 4), dtype=torch.float32, device=dev)
        colors = torch.empty((max(cnt, 1), 4), dtype=torch.float32, device=dev)
        keys = torch.empty((max(cnt, 1),), dtype=torch.int64, device=dev)
        if cnt:
            _check(lib().osl_extract_voxels(self._h, D, centers.data_ptr(), colors.data_ptr(), keys.data_ptr(), cnt,
                                            C.byref(n), None), "osl_extract_voxels")
        return centers[:cnt].cpu().numpy(), colors[:cnt].cpu().numpy(), keys[:cnt].cpu().numpy()


# ---- per-frame image kernels (image_kernels.h:21,24,52) and computeKeys ------------------------------------

def generateVertexMap(depth, fx, fy, device=0):
    torch = _torch()
    h, w = depth.shape
    d = _dev(depth, np.uint16, device)
    out = torch.empty((h * w, 3), dtype=torch.float32, device=d.device)
    _check(lib().osl_generate_vertex_map(d.data_ptr(), out.data_ptr(), w, h, fx, fy, w, h, None),
           "osl_generate_vertex_map")
    return out


def transformVertexMap(points, trans):
    _check(lib().osl_transform_vertex_map(points.data_ptr(), _f(mat_colmajor(trans)), points.shape[0], None),
           "osl_transform_vertex_map")
    return points


def computePointCloudBoundingBox(points, bbox=None):
    b = _f(bbox if bbox is not None else [0.0] * 6)
    _check(lib().osl_point_cloud_bbox(points.data_ptr(), points.shape[0], b, None), "osl_point_cloud_bbox")
    return np.array(list(b), dtype=np.float32)


def computeKeys(points, center, half_edge, max_depth, device=0):
    torch = _torch()
    p = _dev(points, np.float32, device)
    n, stride = p.shape
    keys = torch.empty((n,), dtype=torch.int64, device=p.device)
    _check(lib().osl_compute_keys(p.data_ptr(), stride, n, _f(center), float(half_edge), int(max_depth),
                                  keys.data_ptr(), None), "osl_compute_keys")
    return keys.cpu().numpy()


def meshToVoxelGrid(vertices, triangles, tri_colors4, center, half_edge, max_depth, device=0, want_keys=False):
    """voxelization::meshToVoxelGrid (voxelization.h:21), sparse: -> (centers4, colors4) CUDA tensors [n, 4] in
    Morton-key order (+ int64 keys and int32 lowest-triangle indices when want_keys)."""
    torch = _torch()
    V = _dev(np.asarray(vertices, dtype=np.float32), np.float32, device)
    T = _dev(np.asarray(triangles, dtype=np.int32), np.int32, device)
    Cc = _dev(np.asarray(tri_colors4, dtype=np.float32), np.float32, device) if tri_colors4 is not None else None
    pc, pk, pkeys, ptris, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
    _check(lib().osl_voxelize_mesh(V.data_ptr(), V.shape[0], T.data_ptr(), T.shape[0],
                                   Cc.data_ptr() if Cc is not None else None, _f(center), float(half_edge),
                                   int(max_depth), C.byref(pc), C.byref(pk),
                                   C.byref(pkeys) if want_keys else None, C.byref(ptris) if want_keys else None,
                                   C.byref(n), None), "osl_voxelize_mesh")
    cnt = n.value
    dev = "cuda:%d" % device

    centers = torch.empty((cnt, 4), dtype=torch.float32, device=dev)
    colors = torch.empty((cnt, 4), dtype=torch.float32, device=dev)
    outs = [(pc, centers), (pk, colors)]
    keys = tris = None
    if want_keys:
        keys = torch.empty((cnt,), dtype=torch.int64, device=dev)
        tris = torch.empty((cnt,), dtype=torch.int32, device=dev)
        outs += [(pkeys, keys), (ptris, tris)]
    for ptr, dst in outs:
        if cnt and ptr.value:
            _check(lib().osl_copy_device(dst.data_ptr(), ptr, dst.numel() * dst.element_size()), "osl_copy_device")
        if ptr.value:
            lib().osl_free_device(ptr)
    return (centers, colors, keys, tris) if want_keys else (centers, colors)


def meshToVoxelGridThin(vertices, triangles, tri_colors4, bbox0, bbox1, log_n=8, cube=None, device=0):
    """voxelization::meshToVoxelGrid with the REFERENCE's rule (voxelpipe THIN_RASTER on the dense 2^log_n grid over the
    mesh bounding box, voxelization.cu:24,281-285; osl_voxelize_thin).  cube = (center, half_edge, depth) of the octree
    the grid is meant for orders the voxels by their Morton keys there.  -> (centers4, colors4, cells int32[n,3],
    tris int32[n]) CUDA tensors."""
    torch = _torch()
    V = _dev(np.asarray(vertices, dtype=np.float32), np.float32, device)
    T = _dev(np.asarray(triangles, dtype=np.int32), np.int32, device)
    Cc = _dev(np.asarray(tri_colors4, dtype=np.float32), np.float32, device) if tri_colors4 is not None else None
    pc, pk, pcell, ptri, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
    center, half, depth = cube if cube is not None else ((0.0, 0.0, 0.0), 1.0, 0)
    _check(lib().osl_voxelize_thin(V.data_ptr(), V.shape[0], T.data_ptr(), T.shape[0],
                                   Cc.data_ptr() if Cc is not None else None, _f(bbox0), _f(bbox1), int(log_n),
                                   _f(center), float(half), int(depth), C.byref(pc), C.byref(pk), C.byref(pcell),
                                   C.byref(ptri), C.byref(n), None), "osl_voxelize_thin")
    cnt = n.value
    dev = "cuda:%d" % device
    outs = [(pc, torch.empty((cnt, 4), dtype=torch.float32, device=dev)),
            (pk, torch.empty((cnt, 4), dtype=torch.float32, device=dev)),
            (pcell, torch.empty((cnt, 3), dtype=torch.int32, device=dev)),
            (ptri, torch.empty((cnt,), dtype=torch.int32, device=dev))]
    for ptr, dst in outs:
        if cnt and ptr.value:
            _check(lib().osl_copy_device(dst.data_ptr(), ptr, dst.numel() * dst.element_size()), "osl_copy_device")
        if ptr.value:
            lib().osl_free_device(ptr)
    return tuple(t for _, t in outs)


# ---- reference-shaped classes ------------------------------------------------------------------------------

class BoundingBox:
    """common_types.h:8-17"""

    def __init__(self, bbox0=(0, 0, 0), bbox1=(0, 0, 0)):
        self.bbox0 = np.array(bbox0, dtype=np.float32)
        self.bbox1 = np.array(bbox1, dtype=np.float32)

    def contains(self, other):
        return bool(np.all(self.bbox0 <= other.bbox0) and np.all(self.bbox1 >= other.bbox1))


class Octree:
    """world::Octree (octree.h:83-126).  The root sub-tree is always GPU resident."""

    def __init__(self, resolution, center, size, device=0, reserve_nodes=0):
        self.resolution_ = float(np.float32(resolution))
        self.center_ = tuple(float(c) for c in center)
        self.size_ = float(np.float32(size))
        self.device = device
        self._reserve = reserve_nodes
        self._svo = None

    def _max_depth(self, resolution):
        # octree.cpp:283-284 for node_depth = 0; the unqualified log() is the double overload under g++
        edge_length = np.float32(self.size_) / np.float32(math.pow(2.0, 0.0))
        return int(math.ceil(math.log(float(np.float32(edge_length / np.float32(resolution)))) /
                             float(np.float32(math.log(2.0)))))

    def _tree(self, max_depth):
        if self._svo is None:
            self._svo = SVO(self.center_, self.size_, max_depth, self._reserve, self.device)
        assert self._svo.max_depth == max_depth
        return self._svo

    def addCloud(self, origin, points, colors, size=None, bbox=None):
        """octree.cpp:269-291"""
        self._tree(self._max_depth(self.resolution_)).integrate_points(points, colors)

    def addDepthFrame(self, depth, rgb, fx, fy, pose=IDENTITY):
        """fused main.cpp:39-44 (generateVertexMap + transformVertexMap + addCloud)"""
        self._tree(self._max_depth(self.resolution_)).integrate_depth(depth, rgb, fx, fy, pose)

    def addVoxelGrid(self, centers4, colors4):
        """octree.cpp:293-313"""
        self._tree(self._max_depth(self.resolution_)).integrate_voxels(centers4, colors4)

    def extractVoxelGrid(self, scale):
        """octree.cpp:315-337: max_depth derives from the requested voxel scale"""
        return self._svo.extract_voxels(self._max_depth(scale))

    def extractSVO(self, bbox=None):
        """octree.cpp:339-360 -> (device pointer, n_nodes, center, half size)"""
        return self._svo.view()

    def expandBySize(self, add_size):
        """octree.cpp:362-378 made to work on the GPU tree (Q10): enough doublings for size_ + add_size."""
        ratio = (np.float32(self.size_) + np.float32(add_size)) / np.float32(self.size_)
        layers = int(math.ceil(math.log2(float(ratio)))) if ratio > 1.0 else 0
        if layers < 1:
            return
        if self._svo is not None:
            self._svo.expand(layers)
        self.size_ = float(np.float32(self.size_) * np.float32(2.0 ** layers))

    def boundingBox(self):
        c = np.array(self.center_, dtype=np.float32)
        return BoundingBox(c - np.float32(self.size_), c + np.float32(self.size_))

    @property
    def svo(self):
        return self._svo


class Scene:
    """world::Scene, hot-path methods only (scene.h:35-53)."""

    def __init__(self, device=0):
        self.tree_ = None
        self.device = device

    def addPointCloudToOctree(self, origin, points, colors, size, bbox):
        """scene.cpp:98-113: the first cloud creates Octree(0.01, bbox mid, bbox.bbox1.x) (quirk Q10)."""
        if self.tree_ is None:
            self.tree_ = Octree(0.01, (bbox.bbox1 + bbox.bbox0) / np.float32(2.0), bbox.bbox1[0], self.device)
        self.tree_.addCloud(origin, points, colors, size, bbox)

    def svo(self, bbox=None):
        return self.tree_.extractSVO(bbox)


def coneTraceSVO(svo, resolution, fov, cameraPose, mode=0):
    """rendering::coneTraceSVO(pos, resolution, fov, cameraPose, octree) (cone_tracing_kernels.cu:157)."""
    w, h = int(resolution[0]), int(resolution[1])
    return svo.raycast(w, h, fov, cameraPose, mode)
