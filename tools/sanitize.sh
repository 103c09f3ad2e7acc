#!/bin/sh
# compute-sanitizer over a small end-to-end run (strict + pipelined + host frames + raycast + extraction + map growth
# + camera tracking).
# Usage (on the GPU box): sh tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
tool=${1:-memcheck}
compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
P = g.load_package()
import torch
D, w, h = 12, 160, 120
center, half = P.synth.tree_params(D)
fx, fy = P.synth.focal(w, h)
frames = []
for k in range(6):
    pose = P.synth.orbit_pose(9 * k)
    d, c = P.synth.make_frame(w, h, pose, seed=k)
    frames.append((d, c, torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), pose))
torch.cuda.synchronize()
a = P.SVO(center, half, D)                       # strict
b = P.SVO(center, half, D).set_pipeline(True)    # pipelined, resident inputs
c = P.SVO(center, half, D)                       # pipelined, host frames
for d_, c_, dd, cc, pose in frames:
    a.integrate_depth(dd, cc, fx, fy, pose)
    b.integrate_depth(dd, cc, fx, fy, pose)
    c.integrate_depth_host(d_, c_, fx, fy, pose)
    a.sync()  # feeds the hints -> bucket sort from frame 3 on
pa, pb, pc = a.pool(), b.pool(), c.pool()
assert np.array_equal(pa, pb) and np.array_equal(pa, pc)
img = a.raycast(w, h, 45.0, np.diag([-1.0, 1.0, -1.0, 1.0]).astype(np.float32))
cen, col, keys = a.extract_voxels(D)
pts = np.random.default_rng(0).uniform(-1, 1, size=(5000, 3)).astype(np.float32)
rgb = np.zeros((5000, 3), dtype=np.uint8)
v = P.SVO((0, 0, 0), 1.0, 7)
v.integrate_points(pts, rgb)
cent = np.ones((3000, 4), dtype=np.float32); cent[:, :3] = pts[:3000]
v.integrate_voxels(cent, np.ones((3000, 4), dtype=np.float32) * 0.5)
# map growth, then more frames; camera tracking (both modes) and the free sensor functions
a.expand(1)
a.integrate_depth(frames[0][2], frames[0][3], fx, fy, frames[0][4])
a.sync()
for exact in (False, True):
    cam = P.RGBDCamera(w, h, (fx, fy), exact_jacobian=exact)
    for d_, _, dd, _, _ in frames[:3]:
        cam.update(dd if exact else d_)
    pose = cam.pose()
# tracked SLAM frames: the pose stays on the device (strict and pipelined)
for piped in (False, True):
    cam = P.RGBDCamera(w, h, (fx, fy), exact_jacobian=True)
    tsvo = P.SVO(center, half, D).set_pipeline(piped)
    for _, _, dd, cc, _ in frames[:4]:
        tsvo.integrate_depth_tracked(dd, cc, fx, fy, cam)
    tsvo.sync()
f = P.sensor.bilateralFilter(frames[0][0])
s = P.sensor.subsampleDepth(f)
vm = P.generateVertexMap(f, fx, fy)
nm = P.sensor.generateNormalMap(vm, w, h)
A, b_, pairs = P.sensor.computeICPCost2(vm, nm, vm, nm)
# inputs of > 2^20 keys: k_emit_grid, k_sort_big (keys only and with payload), k_structure_big, leaf-balanced k_levels
rng = np.random.default_rng(1)
Db = 8
cells = np.unique(rng.integers(0, 8 ** Db, size=1_400_000))
ix, iy, iz = cells % 256, (cells // 256) % 256, cells // 65536
big = np.ones((cells.size, 4), dtype=np.float32)
big[:, 0], big[:, 1], big[:, 2] = (ix + 0.5) / 128.0 - 1.0, (iy + 0.5) / 128.0 - 1.0, (iz + 0.5) / 128.0 - 1.0
bigcol = rng.uniform(0, 1, size=big.shape).astype(np.float32)
g1, g2 = P.SVO((0, 0, 0), 1.0, Db, reserve_nodes=1 << 23), P.SVO((0, 0, 0), 1.0, Db, reserve_nodes=1 << 23)
gb, gc = torch.from_numpy(big).cuda(), torch.from_numpy(bigcol).cuda()
order = torch.from_numpy(np.argsort(P.computeKeys(big[:, :3], (0, 0, 0), 1.0, Db), kind="stable")).cuda()
for _ in range(3):
    g1.integrate_voxels(gb, gc)                  # shuffled: the sort runs
    g2.integrate_voxels(gb[order].contiguous(), gc[order].contiguous())  # Morton order: skipped
assert g1.size == g2.size and np.array_equal(g1.pool()[0::2], g2.pool()[0::2])
g3 = P.SVO((0, 0, 0), 1.0, Db, reserve_nodes=1 << 23)
g4 = P.SVO((0, 0, 0), 1.0, Db, reserve_nodes=1 << 23)
g3.integrate_points(big[:, :3], np.zeros((big.shape[0], 3), dtype=np.uint8))   # big point cloud: pair sort
g4.integrate_voxels(gb, gc)
assert g3.size == g4.size and np.array_equal(g3.pool()[0::2], g4.pool()[0::2])
# voxel keys on / next to cell boundaries (k_grid_fix, k_grid_fix_check) and an invalid voxel (k_grid_compact)
edge = big[:200000].copy()
edge[::3, 0] = np.float32(-1.0) + np.float32(2.0 / 256) * np.arange(edge[::3].shape[0], dtype=np.float32) % np.float32(2.0)
edge[1::3, 1] = np.nextafter(edge[1::3, 1] - np.float32(1.0 / 256), np.float32(9))
edge[77, 2] = np.inf
g5 = P.SVO((0, 0, 0), 1.0, Db, reserve_nodes=1 << 22)
g5.integrate_voxels(edge, bigcol[:200000])
# the sparse mesh voxeliser (append + sort + unique) and the reference's thin rule
Vm, Tm = P.synth.icosphere(3, 0.7, (0.05, 0.0, -0.03))
cm, km = P.meshToVoxelGrid(Vm, Tm, None, (0, 0, 0), 1.0, 8)
_, _, cells, tris = P.meshToVoxelGridThin(Vm, Tm, None, Vm.min(axis=0), Vm.max(axis=0), 6)
print("sanitize run ok:", a.size, img.shape, keys.size, v.size, pairs, g1.size, g5.size, cm.shape[0], cells.shape[0])
PY
