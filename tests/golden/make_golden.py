"""Generates the golden vectors under tests/golden/ by running the REFERENCE's own CUDA code
(oracle/_ref/libosl_ref*.so = /root/reference sources compiled unmodified for sm_100a + oracle/ref_shim.cu) on a GPU.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   then copy gpurun_out/golden/*.npz here.

Inputs are seeded and duplicate-free so the reference is deterministic (its duplicate-key and node-0 races, quirks
Q7/Q6, are avoided or masked: golden pools store word index 1 -- node 0's value -- as produced, the tests skip it
whenever more than one warp wrote it)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import LOOK_PLUS_Z, random_pose, unique_voxel_points, view_for_pose  # noqa: E402
from oracle import ref as R  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)

    # G1: the hand-derived KAT of SURVEY.md section 8c, confirmed by the real reference
    pts = np.array([[0.3, 0.3, 0.3], [0.9, 0.9, 0.9], [-0.9, 0.2, 0.6]], dtype=np.float32)
    rgb = np.array([[200, 100, 50], [10, 20, 30], [255, 255, 255]], dtype=np.uint8)
    t = R.RefSVO((0, 0, 0), 1.0, 2)
    t.integrate_points(pts, rgb)
    p1 = t.pool()
    t.integrate_points(pts, rgb)
    p2 = t.pool()
    np.savez_compressed(os.path.join(out_dir, "g1_kat.npz"), pts=pts, rgb=rgb, pool1=p1, pool2=p2)

    # G2: duplicate-free clouds, three frames, D = 8 (unmodified reference)
    rng = np.random.default_rng(1234)
    center, half, D = (0.05, -0.1, 0.2), 1.0, 8
    t = R.RefSVO(center, half, D)
    frames = {}
    for f in range(3):
        pts = unique_voxel_points(rng, 2500, center, half, D)
        rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
        t.integrate_points(pts, rgb)
        frames["pts%d" % f] = pts
        frames["rgb%d" % f] = rgb
        frames["pool%d" % f] = t.pool()
    np.savez_compressed(os.path.join(out_dir, "g2_clouds_d8.npz"), center=np.array(center, dtype=np.float32),
                        half=np.float32(half), D=D, **frames)

    # G3: generateVertexMap + transformVertexMap
    rng = np.random.default_rng(77)
    w, h = 64, 48
    depth = rng.integers(0, 16000, size=(h, w)).astype(np.uint16)
    pose = random_pose(rng)
    fx, fy = np.float32(532.57 * w / 640), np.float32(531.54 * h / 480)
    np.savez_compressed(os.path.join(out_dir, "g3_vertex_map.npz"), depth=depth, pose=pose, fx=fx, fy=fy,
                        xyz=R.vertex_map(depth, fx, fy, None), xyz_t=R.vertex_map(depth, fx, fy, pose))

    # G4: svoFromVoxelGrid (float colours incl. 1.0 -> Q15; colours scrambled by the key sort -> Q11)
    rng = np.random.default_rng(4321)
    D = 6
    pts = unique_voxel_points(rng, 2000, (0, 0, 0), 1.0, D)
    centers = np.ones((pts.shape[0], 4), dtype=np.float32)
    centers[:, :3] = pts
    colors = rng.uniform(0, 1, size=centers.shape).astype(np.float32)
    colors[::7] = 1.0
    t = R.RefSVO((0, 0, 0), 1.0, D)
    t.integrate_voxels(centers, colors)
    pa = t.pool()
    t.integrate_voxels(centers, colors)
    np.savez_compressed(os.path.join(out_dir, "g4_voxels_d6.npz"), centers=centers, colors=colors, D=D, pool1=pa,
                        pool2=t.pool())

    # G5: raycast + extraction of a saturated tree (66 inserts of one duplicate-free cloud)
    rng = np.random.default_rng(555)
    D = 6
    center, half = (0.0, 0.0, 0.0), 1.28
    pts = unique_voxel_points(rng, 6000, center, half, D)
    pts = pts[(np.abs(np.linalg.norm(pts - np.array([0, 0, 0.7]), axis=1) - 0.45) < 0.05) |
              (np.abs(pts[:, 2] - 1.2) < 0.03)]
    rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
    t = R.RefSVO(center, half, D)
    for _ in range(66):
        t.integrate_points(pts, rgb)
    pool = t.pool()
    views = [LOOK_PLUS_Z, view_for_pose(random_pose(rng, 0.15, 0.25)), np.eye(4, dtype=np.float32)]
    imgs = [t.raycast(64, 48, 45.0, v)[0] for v in views]
    cen, col = t.extract_voxels(D)
    np.savez_compressed(os.path.join(out_dir, "g5_raycast_extract.npz"), pts=pts, rgb=rgb, D=D,
                        center=np.array(center, dtype=np.float32), half=np.float32(half), pool=pool,
                        views=np.stack(views), imgs=np.stack(imgs), ex_centers=cen, ex_colors=col)

    # G6: deep tree (D = 12) through the "ref + 64-bit patch" build
    if R.available(True):
        rng = np.random.default_rng(66)
        D = 12
        half = np.float32(0.01)
        for _ in range(D):
            half = np.float32(half * np.float32(2))
        t = R.RefSVO((0, 0, 0), float(half), D, patched64=True)
        frames = {}
        for f in range(2):
            pts = unique_voxel_points(rng, 2000, (0, 0, 0), float(half), D)
            rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
            t.integrate_points(pts, rgb)
            frames["pts%d" % f] = pts
            frames["rgb%d" % f] = rgb
            frames["pool%d" % f] = t.pool()
        np.savez_compressed(os.path.join(out_dir, "g6_clouds_d12_ref64.npz"), half=half, D=D, **frames)
    # G7: camera tracking -- the reference's image / localization kernels and its own RGBDCamera::update
    # (image_kernels.cu:104-321, localization_kernels.cu, rgbd_camera.cpp) on four frames of a small synthetic orbit
    import __graft_entry__ as graft
    synth = graft.load_package().synth
    w, h = 160, 120
    fx, fy = synth.focal(w, h)
    depths = []
    for k in range(4):
        M = np.eye(4)
        a = np.radians(0.3 * k)
        M[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        M[:3, 3] = [0.01 * k, 0.0, 0.005 * k]
        depths.append(synth.make_frame(w, h, M.astype(np.float32), seed=k, invalid_frac=0.02, noise_mm=2)[0])
    depths = np.stack(depths)
    filt = [R.bilateral(d) for d in depths[:2]]
    sub = R.subsample_depth(filt[0])
    maps = []
    for f in filt:
        v = R.vertex_map(f, fx, fy, None)
        maps.append((v, R.normal_map(v, w, h)))
    A, b = R.icp_cost2(maps[0][0], maps[0][1], maps[1][0], maps[1][1], w, h)
    rng = np.random.default_rng(7)
    rgb = rng.integers(0, 256, size=(500, 3), dtype=np.uint8)
    cam = R.RefTracker(w, h, fx, fy)
    ori, pos = [], []
    for d in depths:
        cam.update(d)
        ori.append(cam.orientation())
        pos.append(cam.position())
    np.savez_compressed(os.path.join(out_dir, "g7_tracking.npz"), depths=depths, fx=fx, fy=fy, bilateral0=filt[0],
                        subsample0=sub, normals0=maps[0][1], icp_A=A, icp_b=b, rgb=rgb,
                        intensity=R.color_to_intensity(rgb), orientation=np.stack(ori), position=np.stack(pos))
    print("golden vectors written to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__)))
