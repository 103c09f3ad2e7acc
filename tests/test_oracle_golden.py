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
