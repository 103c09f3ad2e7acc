"""CPU-only: pins the CPU oracle against golden vectors produced by the REFERENCE's own CUDA code on a B200
(tests/golden/make_golden.py; the reference ships no tests or fixtures of its own)."""
import os

import numpy as np
import pytest

from common import float_bits_equal
from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("%s not generated yet (run tests/golden/make_golden.py on a GPU box)" % name)
    return np.load(path)


def _eq_except_node0_value(a, b):
    a, b = a.copy(), b.copy()
    a[1] = b[1] = 0  # Q6: racy in the reference when more than one warp writes it
    return np.array_equal(a, b)


def test_g1_kat_confirmed_by_the_reference():
    g = _load("g1_kat.npz")
    t = orc.OracleSVO((0, 0, 0), 1.0, 2)
    t.integrate_points(g["pts"], g["rgb"])
    assert np.array_equal(t.pool(), g["pool1"])  # 3 keys = one warp: even node 0's value is deterministic
    t.integrate_points(g["pts"], g["rgb"])
    assert np.array_equal(t.pool(), g["pool2"])


def test_g2_duplicate_free_clouds_d8():
    g = _load("g2_clouds_d8.npz")
    t = orc.OracleSVO(tuple(g["center"]), float(g["half"]), int(g["D"]))
    for f in range(3):
        t.integrate_points(g["pts%d" % f], g["rgb%d" % f])
        assert t.size * 2 == g["pool%d" % f].size
        assert _eq_except_node0_value(t.pool(), g["pool%d" % f]), "frame %d" % f


def test_g3_vertex_map_and_transform():
    g = _load("g3_vertex_map.npz")
    xyz = orc.vertex_map(g["depth"], float(g["fx"]), float(g["fy"]))
    assert float_bits_equal(xyz, g["xyz"])
    assert float_bits_equal(orc.transform(xyz, g["pose"]), g["xyz_t"])


def test_g4_voxel_grid_d6():
    g = _load("g4_voxels_d6.npz")
    t = orc.OracleSVO((0, 0, 0), 1.0, int(g["D"]))
    t.integrate_voxels(g["centers"], g["colors"])
    assert _eq_except_node0_value(t.pool(), g["pool1"])
    t.integrate_voxels(g["centers"], g["colors"])
    assert _eq_except_node0_value(t.pool(), g["pool2"])


def test_g5_raycast_and_extraction():
    g = _load("g5_raycast_extract.npz")
    center, half, D = tuple(g["center"]), float(g["half"]), int(g["D"])
    t = orc.OracleSVO(center, half, D)
    for _ in range(66):
        t.integrate_points(g["pts"], g["rgb"])
    assert _eq_except_node0_value(t.pool(), g["pool"])
    t.load(g["pool"])  # identical pool (incl. the racy word) for the image comparison
    for view, want in zip(g["views"], g["imgs"]):
        img = t.raycast(64, 48, 45.0, view)
        assert np.array_equal(img, want), "%d pixels differ" % np.count_nonzero(np.any(img != want, axis=2))
    c, k, _ = t.extract_voxels(D)
    assert float_bits_equal(c, g["ex_centers"]) and float_bits_equal(k, g["ex_colors"])


def test_g6_deep_tree_ref64():
    g = _load("g6_clouds_d12_ref64.npz")
    t = orc.OracleSVO((0, 0, 0), float(g["half"]), int(g["D"]))
    for f in range(2):
        t.integrate_points(g["pts%d" % f], g["rgb%d" % f])
        assert _eq_except_node0_value(t.pool(), g["pool%d" % f])


def test_g7_camera_tracking_against_the_reference_kernels_and_rgbd_camera():
    """bilateralFilter, subsampleDepth, generateNormalMap, computeICPCost2, colorToIntensity and four frames of the
    reference's own RGBDCamera::update, all run by the reference's code on a B200."""
    g = _load("g7_tracking.npz")
    depths, fx, fy = g["depths"], float(g["fx"]), float(g["fy"])
    h, w = depths.shape[1:]
    f0 = orc.bilateral(depths[0])
    diff = np.abs(f0.astype(np.int64) - g["bilateral0"].astype(np.int64))
    assert diff.max() <= 1 and np.mean(diff != 0) <= 0.005   # exp2f on the CPU vs MUFU.EX2 (see osl_oracle_track.c)
    # downstream of the reference's OWN filtered image everything is exact or a float sum
    assert np.array_equal(orc.subsample_depth(g["bilateral0"]), g["subsample0"])
    v0 = orc.vertex_map(g["bilateral0"], fx, fy)
    n0 = orc.normal_map(v0, w, h)
    assert float_bits_equal(n0, g["normals0"])
    f1 = orc.bilateral(depths[1])
    v1 = orc.vertex_map(f1, fx, fy)
    n1 = orc.normal_map(v1, w, h)
    A, b, pairs = orc.icp_cost(v0, n0, v1, n1)
    assert pairs > 0.3 * w * h
    assert np.abs(A - g["icp_A"]).max() <= 2e-4 * np.abs(g["icp_A"]).max()   # incl. the 1 mm bilateral differences
    assert np.abs(b - g["icp_b"]).max() <= 2e-4 * np.abs(g["icp_A"]).max()
    assert float_bits_equal(orc.color_to_intensity(g["rgb"]), g["intensity"])
    t = orc.OracleTracker(w, h, fx, fy)
    for k in range(depths.shape[0]):
        t.update(depths[k])
        assert np.abs(t.orientation() - g["orientation"][k]).max() <= 1e-4, k
        assert np.abs(t.position() - g["position"][k]).max() <= 1e-4
    assert not np.array_equal(g["orientation"][-1], np.eye(3, dtype=np.float32))
    assert np.array_equal(g["position"][-1], np.zeros(3, dtype=np.float32))   # quirk Q18, in the reference itself
