"""The C++ host side above the C ABI (octree-slam_b200/host): libosl_host.so re-creates the reference's
world / rendering / sensor interface; osl_main replays main.cpp:31-62 headless.  CPU part: the library exports the
reference's function and class symbols.  GPU part: the replay's pool and image equal the oracle's."""
import math
import os
import struct
import subprocess

import numpy as np
import pytest

from common import LOOK_PLUS_Z, ROOT, pkg
from oracle import oracle as orc

PKG_DIR = os.path.join(ROOT, "octree-slam_b200")
HOST_LIB = os.path.join(PKG_DIR, "libosl_host.so")
HOST_MAIN = os.path.join(PKG_DIR, "osl_main")


def test_host_library_exports_the_reference_interface():
    pkg()  # build() has produced the libraries
    assert os.path.exists(HOST_LIB) and os.path.exists(HOST_MAIN)
    syms = subprocess.check_output(["nm", "-D", "--demangle", "--defined-only", HOST_LIB], text=True)
    # the seam functions have C linkage INSIDE their namespaces, exactly like the reference's declarations
    # (svo.h:14-18, cone_tracing_kernels.h:16, image_kernels.h:21-55, localization_kernels.h:36-42,
    # voxelization.h:21): unmangled symbols, so objects compiled against the reference's headers link here
    plain = set(subprocess.check_output(["nm", "-D", "--defined-only", HOST_LIB], text=True).split())
    for name in ["svoFromPointCloud", "svoFromVoxelGrid", "extractVoxelGridFromSVO", "coneTraceSVO",
                 "generateVertexMap", "transformVertexMap", "computePointCloudBoundingBox", "generateNormalMap",
                 "bilateralFilter", "colorToIntensity", "transformNormalMap", "computeICPCost2", "computeICPCost",
                 "computeRGBDCost", "meshToVoxelGrid"]:
        assert name in plain, "libosl_host.so does not define the C-linkage symbol %s" % name
    for name in ["octree_slam::world::Octree::addCloud(",
                 "octree_slam::world::Octree::addVoxelGrid(", "octree_slam::world::Octree::extractVoxelGrid(",
                 "octree_slam::world::Octree::extractSVO(", "octree_slam::world::Scene::addPointCloudToOctree(",
                 "octree_slam::world::Scene::extractVoxelGridFromOctree(",
                 "octree_slam::rendering::CUDARenderer::coneTraceSVO(",
                 "octree_slam::world::Octree::expandBySize(", "BoundingBox::contains(", "BoundingBox::distanceOutside(",
                 "void octree_slam::sensor::subsampleDepth<unsigned short>(", "void octree_slam::sensor::subsample<float>(",
                 "octree_slam::sensor::ICPFrame::ICPFrame(",
                 "octree_slam::sensor::RGBDCamera::update(", "octree_slam::sensor::RGBDCamera::position(",
                 "octree_slam::sensor::RGBDCamera::orientation(", "octree_slam::sensor::RGBDCamera::camera("]:
        assert name in syms, "libosl_host.so does not define %s" % name
    # the host side contains no device code and needs only the C ABI + cudart
    needed = subprocess.check_output(["readelf", "-d", HOST_LIB], text=True)
    assert "libosl_b200.so" in needed and "libcudart" in needed


REF_GLM = "/root/reference/external/include"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_GLM, "glm")),
                    reason="the reference tree (its vendored glm 0.9.5.4) is not mounted here")
def test_shim_builds_against_the_reference_glm(tmp_path):
    """a maintainer compiles the shim inside the reference tree, i.e. with the reference's
    own glm first on the include path; the replay of main.cpp must compile unchanged too."""
    inc = ["-I" + REF_GLM, "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]
    host = os.path.join(PKG_DIR, "host")
    out = str(tmp_path / "libosl_host_refglm.so")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-fPIC", "-w", "-shared"] + inc +
                          [os.path.join(host, "osl_host.cpp"), "-o", out], stderr=subprocess.DEVNULL)
    syms = subprocess.check_output(["nm", "-D", "--demangle", "--defined-only", out], text=True)
    assert "glm::detail::tvec3<float" in syms            # really the reference's glm types in the signatures
    assert "octree_slam::sensor::RGBDCamera::update(" in syms and "octree_slam::world::Octree::addCloud(" in syms
    subprocess.check_call(["g++", "-std=c++14", "-w", "-fsyntax-only"] + inc + [os.path.join(host, "osl_main.cpp")],
                          stderr=subprocess.DEVNULL)


DROPIN = os.path.join(ROOT, "oracle", "_ref", "ref_octree_dropin")


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/world/octree.cpp"),
                    reason="the reference tree is not mounted here")
def test_reference_octree_cpp_links_against_the_shim():
    """Link-level drop-in: the reference's OWN src/world/octree.cpp, compiled against the
    reference's OWN headers and glm, links against libosl_host.so -- its undefined seam symbols are the unmangled names
    of svo.h:14-18 and libosl_host.so defines exactly those."""
    pkg()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"])
    obj = os.path.join(ROOT, "oracle", "_ref", "obj", "ref_octree.o")
    undefined = set(subprocess.check_output(["nm", "-u", obj], text=True).split())
    defined = set(subprocess.check_output(["nm", "-D", "--defined-only", HOST_LIB], text=True).split())
    for name in ("svoFromPointCloud", "svoFromVoxelGrid", "extractVoxelGridFromSVO"):
        assert name in undefined and name in defined
    needed = subprocess.check_output(["readelf", "-d", DROPIN], text=True)
    assert "libosl_host.so" in needed  # (which in turn needs libosl_b200.so, checked above)
    # nothing of the reference's device code is in the binary: the only CUDA it can reach is ours
    syms = subprocess.check_output(["nm", "--demangle", DROPIN], text=True)
    assert "splitNodes" not in syms and "coneTrace(" not in syms and "fillNodes" not in syms


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(DROPIN), reason="oracle/_ref/ref_octree_dropin not built (needs /root/reference)")
def test_reference_octree_class_runs_on_our_library(tmp_path):
    """The reference's own world::Octree (its octree.cpp object) drives our library through the unmangled seam:
    addCloud x 2 -> extractSVO -> coneTraceSVO -> extractVoxelGrid; pool, image and voxels equal the oracle's."""
    P = pkg()
    rng = np.random.default_rng(77)
    D, w, h = 7, 96, 72
    center, half = (0.0, 0.0, 0.0), P.synth.tree_params(D)[1]
    pts = rng.uniform(-0.9 * half, 0.9 * half, size=(20000, 3)).astype(np.float32)
    pts[::50] = np.float32(np.inf)  # invalid points (svo.cu:38)
    rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
    view = LOOK_PLUS_Z
    src, dst = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(src, "wb") as f:
        f.write(struct.pack("<iii", pts.shape[0], w, h))
        f.write(struct.pack("<3f", *center))
        f.write(struct.pack("<fff", half, 0.01, 45.0))
        f.write(np.ascontiguousarray(view.T, dtype=np.float32).tobytes())  # column-major
        f.write(pts.tobytes())
        f.write(rgb.tobytes())
    out = subprocess.run([DROPIN, src, dst], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = open(dst, "rb").read()
    n_nodes = struct.unpack_from("<i", raw, 0)[0]
    pool = np.frombuffer(raw, dtype=np.uint32, count=2 * n_nodes, offset=4)
    off = 4 + 8 * n_nodes
    img = np.frombuffer(raw, dtype=np.uint8, count=w * h * 4, offset=off).reshape(h, w, 4)
    off += w * h * 4
    n_vox = struct.unpack_from("<i", raw, off)[0]
    cen = np.frombuffer(raw, dtype=np.float32, count=4 * n_vox, offset=off + 4).reshape(n_vox, 4)
    col = np.frombuffer(raw, dtype=np.float32, count=4 * n_vox, offset=off + 4 + 16 * n_vox).reshape(n_vox, 4)
    ref = orc.OracleSVO(center, half, D)   # Octree::addCloud derives max_depth = 7 from size / resolution (octree.cpp:284)
    ref.integrate_points(pts, rgb)
    ref.integrate_points(pts, rgb)
    assert n_nodes == ref.size
    assert np.array_equal(pool, ref.pool())
    assert np.array_equal(img, ref.raycast(w, h, 45.0, view))
    rc, rk, _ = ref.extract_voxels(D)
    assert n_vox == rc.shape[0] and np.array_equal(cen, rc) and np.array_equal(col, rk)


def _write_frames(path, w, h, frames, fx, fy):
    with open(path, "wb") as f:
        f.write(struct.pack("<iiiff", w, h, len(frames), fx, fy))
        for pose, depth, rgb in frames:
            f.write(np.ascontiguousarray(np.asarray(pose, dtype=np.float32).T).tobytes())  # column-major
            f.write(np.ascontiguousarray(depth, dtype=np.uint16).tobytes())
            f.write(np.ascontiguousarray(rgb, dtype=np.uint8).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_main_loop_replay_matches_oracle(tmp_path, fused):
    P = pkg()
    w, h, n = 160, 120, 3
    fx, fy = P.synth.focal(w, h)
    frames = []
    for k in range(n):
        pose = P.synth.orbit_pose(20 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        frames.append((pose, depth, rgb))
    fpath = str(tmp_path / "frames.bin")
    _write_frames(fpath, w, h, frames, fx, fy)
    out = str(tmp_path / "out")
    log = subprocess.check_output([HOST_MAIN, fpath, out] + (["fused"] if fused else []), text=True, timeout=120)
    assert "osl_main:" in log

    raw = open(out + ".pool", "rb").read()
    n_nodes, cx, cy, cz, half = struct.unpack("<iffff", raw[:20])
    pool = np.frombuffer(raw[20:], dtype=np.uint32)
    assert pool.size == 2 * n_nodes
    img = np.frombuffer(open(out + ".rgba", "rb").read(), dtype=np.uint8).reshape(h, w, 4)

    # the same loop on the oracle: Scene::addPointCloudToOctree creates Octree(0.01, bbox mid, bbox.bbox1.x)
    # from the FIRST cloud (scene.cpp:100-102), max_depth from octree.cpp:283-284
    ref = None
    for pose, depth, rgb in frames:
        xyz = orc.transform(orc.vertex_map(depth, fx, fy), pose)
        if ref is None:
            b = orc.bbox(xyz)
            center = (b[3:] + b[:3]) / np.float32(2.0)
            size = float(b[3])
            D = int(math.ceil(math.log(float(np.float32(np.float32(size) / np.float32(0.01)))) /
                              float(np.float32(math.log(2.0)))))
            ref = orc.OracleSVO(tuple(center), size, D)
        ref.integrate_points(xyz, rgb.reshape(-1, 3))
    assert (np.float32(cx), np.float32(cy), np.float32(cz)) == tuple(np.float32(c) for c in ref.center)
    assert np.float32(half) == np.float32(ref.half_edge)
    assert n_nodes == ref.size
    assert np.array_equal(pool, ref.pool())
    assert np.array_equal(img, ref.raycast(w, h, 45.0, LOOK_PLUS_Z))


@pytest.mark.gpu
def test_scene_voxelize_meshes_replay_matches_oracle(tmp_path):
    """Scene::voxelizeMeshes(true) (scene.cpp:64-85) through the C++ shim: mesh -> VoxelGrid -> Octree::addVoxelGrid"""
    P = pkg()
    V, T = P.synth.icosphere(2, 0.6, (0.9, 0.8, 0.7))  # bbox1.x = 1.5 is the half edge the reference would use
    mpath = str(tmp_path / "mesh.bin")
    with open(mpath, "wb") as f:
        f.write(struct.pack("<ii", V.shape[0], T.shape[0]))
        f.write(V.tobytes())
        f.write(T.tobytes())
    out = str(tmp_path / "mesh_out")
    log = subprocess.check_output([HOST_MAIN, "mesh", mpath, out], text=True, timeout=120)
    assert "osl_main: mesh" in log
    raw = open(out + ".pool", "rb").read()
    n_nodes, cx, cy, cz, half = struct.unpack("<iffff", raw[:20])
    pool = np.frombuffer(raw[20:], dtype=np.uint32)
    # the oracle side of the same call sequence: scale = bbox1.x / 256, Octree(scale, bbox mid, bbox1.x) -> depth 8
    b0, b1 = V.min(axis=0), V.max(axis=0)
    center = (b1 + b0) / np.float32(2.0)
    size = float(b1[0])
    assert (np.float32(cx), np.float32(cy), np.float32(cz)) == tuple(np.float32(c) for c in center)
    assert np.float32(half) == np.float32(size)
    # meshToVoxelGrid with the reference's signature = the reference's rule: voxelpipe THIN_RASTER on the dense 256^3
    # grid over the mesh BOUNDING BOX (oracle/osl_oracle_thin.c), centres per getCenterFromIndex (voxelization.cu:58-78),
    # handed to svoFromVoxelGrid in ascending order of their keys in the octree cube (ties by grid index)
    cells, tris = orc.voxelize_thin(V, T, b0, b1, 8)
    f32 = np.float32
    td = (b1 - b0).astype(f32) / f32(32)
    pd = td / f32(8)
    cen = np.ones((cells.shape[0], 4), dtype=f32)
    cen[:, :3] = (b0[None, :] + (cells // 8).astype(f32) * td[None, :] + (cells % 8).astype(f32) * pd[None, :] +
                  (pd / f32(2.0))[None, :]).astype(f32)
    keys = orc.compute_keys(cen[:, :3], tuple(center), size, 8)
    lin = (cells[:, 2].astype(np.int64) * 256 + cells[:, 1]) * 256 + cells[:, 0]
    order = np.lexsort((lin, keys))
    cen = np.ascontiguousarray(cen[order])
    colors = np.zeros((cen.shape[0], 4), dtype=np.float32)
    colors[:, 1] = np.float32(255 / 255.0)  # no texture: green
    ref = orc.OracleSVO(tuple(center), size, 8)
    ref.integrate_voxels(cen, colors)
    assert n_nodes == ref.size
    assert np.array_equal(pool, ref.pool())


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["track", "slam"])
def test_main_loop_with_camera_tracking_matches_oracle(tmp_path, mode):
    """main.cpp:33-44 with line 35 live: sensor::RGBDCamera estimates every pose, the map is built with the estimates"""
    P = pkg()
    w, h, n = 160, 120, 4
    fx, fy = P.synth.focal(w, h)
    frames = []
    for k in range(n):
        pose = np.eye(4, dtype=np.float32)
        pose[0, 3] = 0.01 * k
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k, invalid_frac=0.01, noise_mm=1)
        frames.append((pose, depth, rgb))
    fpath = str(tmp_path / "frames.bin")
    _write_frames(fpath, w, h, frames, fx, fy)
    out = str(tmp_path / "out")
    log = subprocess.check_output([HOST_MAIN, fpath, out, mode], text=True, timeout=120)
    assert "osl_main:" in log
    est = np.frombuffer(open(out + ".poses", "rb").read(), dtype=np.float32).reshape(n, 4, 4).transpose(0, 2, 1)
    track = orc.OracleTracker(w, h, fx, fy, exact_jacobian=True)
    ref = None
    for k, (_, depth, rgb) in enumerate(frames):
        track.update(depth)
        assert np.abs(est[k] - track.pose()).max() <= 1e-4, (k, est[k], track.pose())
        xyz = orc.transform(orc.vertex_map(depth, fx, fy), est[k])   # the map follows the ESTIMATED poses
        if ref is None:
            b = orc.bbox(xyz)
            center = (b[3:] + b[:3]) / np.float32(2.0)
            size = float(b[3])
            D = int(math.ceil(math.log(float(np.float32(np.float32(size) / np.float32(0.01)))) /
                              float(np.float32(math.log(2.0)))))
            ref = orc.OracleSVO(tuple(center), size, D)
        ref.integrate_points(xyz, rgb.reshape(-1, 3))
    assert np.abs(est[-1][0, 3] - 0.03) < 6e-3   # the tracker follows the 3 cm of true motion
    raw = open(out + ".pool", "rb").read()
    n_nodes = struct.unpack("<i", raw[:4])[0]
    assert n_nodes == ref.size
    assert np.array_equal(np.frombuffer(raw[20:], dtype=np.uint32), ref.pool())
