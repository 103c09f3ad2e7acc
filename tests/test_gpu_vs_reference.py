"""GPU tests against the REFERENCE's own CUDA code (oracle/_ref/libosl_ref*.so, built from /root/reference by
oracle/Makefile).  Where the reference is deterministic the comparison is bit-exact on the whole pool / image; the two
places where the reference races (Q6: node 0's value word, Q7: duplicate keys) are compared on what is well defined."""
import numpy as np
import pytest

from common import (check_frame_outcome, float_bits_equal, FLAG, LOOK_PLUS_Z, pkg, random_pose, unique_voxel_points,
                    view_for_pose)
from oracle import oracle as orc
from oracle import ref as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref/libosl_ref.so not built (needs /root/reference)")]


@pytest.fixture(scope="module")
def P():
    return pkg()


def _eq_except_node0_value(a, b):
    a, b = a.copy(), b.copy()
    a[1] = b[1] = 0  # Q6: the reference's node-0 value word is written while other threads still read it
    return np.array_equal(a, b)


@pytest.mark.parametrize("D", [2, 6, 8, 10])
def test_duplicate_free_clouds_match_reference_bit_exact(P, D):
    rng = np.random.default_rng(40 + D)
    n = min(6000, 8 ** D // 3)
    svo = P.SVO((0.05, -0.1, 0.2), 1.0, D)
    ref = R.RefSVO((0.05, -0.1, 0.2), 1.0, D)
    for frame in range(3):
        pts = unique_voxel_points(rng, n, (0.05, -0.1, 0.2), 1.0, D)
        rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
        svo.integrate_points(pts, rgb)
        ref.integrate_points(pts, rgb)
        assert svo.size == ref.size
        assert _eq_except_node0_value(svo.pool(), ref.pool()), "frame %d" % frame


@pytest.mark.parametrize("D", [11, 14, 16])
def test_deep_trees_match_patched_reference(P, D):
    """D >= 11: the unmodified reference truncates keys to 32 bits; compare with 'ref + 64-bit patch'."""
    if not R.available(True):
        pytest.skip("libosl_ref64.so not built")
    rng = np.random.default_rng(D)
    center, half = P.synth.tree_params(D)
    svo = P.SVO(center, half, D)
    ref = R.RefSVO(center, half, D, patched64=True)
    for frame in range(2):
        pts = unique_voxel_points(rng, 5000, center, half, D)
        rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
        svo.integrate_points(pts, rgb)
        ref.integrate_points(pts, rgb)
        assert svo.size == ref.size
        assert _eq_except_node0_value(svo.pool(), ref.pool())


def test_depth_frames_structure_matches_reference(P):
    """Real frames have duplicate keys (Q7, svo.cu:366-381): node INDICES / child pointers must match exactly, and
    EVERY leaf is checked against its legal outcomes after every frame -- ours must be the blend of the lowest input
    index that maps to the leaf (alpha += 2 once), the reference's must be reachable by m >= 1 successive blends of
    inputs of that leaf (alpha += 2m); inner nodes are averageChildren of their tiles on both sides."""
    D, w, h = 8, 320, 240
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D)
    ref = R.RefSVO(center, half, D)
    a0 = b0 = np.zeros(0, dtype=np.uint32)
    for k in range(3):
        pose = P.synth.orbit_pose(30 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        keys = orc.compute_keys(orc.transform(orc.vertex_map(depth, fx, fy), pose), center, half, D)
        svo.integrate_depth(depth, rgb, fx, fy, pose)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
        assert svo.size == ref.size
        a, b = svo.pool(), ref.pool()
        assert np.array_equal(a[0::2], b[0::2]), "child pointers / node indices differ"
        n = check_frame_outcome(a0, a, keys, rgb, D, canonical=True)
        assert check_frame_outcome(b0, b, keys, rgb, D, canonical=False) == n
        # alpha: ours is the canonical (minimum legal) outcome
        assert np.all((a[1::2] >> 24) <= (b[1::2] >> 24))
        a0, b0 = a, b


def test_vertex_map_matches_reference(P):
    rng = np.random.default_rng(9)
    w, h = 160, 120
    depth = rng.integers(0, 16000, size=(h, w)).astype(np.uint16)
    fx, fy = P.synth.focal(w, h)
    pose = random_pose(rng)
    pts = P.transformVertexMap(P.generateVertexMap(depth, fx, fy), pose).cpu().numpy()
    want = R.vertex_map(depth, fx, fy, pose)
    assert float_bits_equal(pts, want)


def test_voxel_grid_matches_reference(P):
    rng = np.random.default_rng(21)
    D = 7
    pts = unique_voxel_points(rng, 5000, (0, 0, 0), 1.0, D)
    centers = np.ones((pts.shape[0], 4), dtype=np.float32)
    centers[:, :3] = pts
    colors = rng.uniform(0, 1, size=centers.shape).astype(np.float32)
    colors[::7] = 1.0
    svo = P.SVO((0, 0, 0), 1.0, D)
    ref = R.RefSVO((0, 0, 0), 1.0, D)
    for _ in range(2):
        svo.integrate_voxels(centers, colors)
        ref.integrate_voxels(centers, colors)
    assert svo.size == ref.size
    assert _eq_except_node0_value(svo.pool(), ref.pool())


def test_extract_matches_reference(P):
    rng = np.random.default_rng(22)
    D = 7
    pts = unique_voxel_points(rng, 4000, (0, 0, 0), 1.0, D)
    rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
    svo = P.SVO((0, 0, 0), 1.0, D)
    svo.integrate_points(pts, rgb)
    ref = R.RefSVO((0, 0, 0), 1.0, D)
    ref.load(svo.pool())
    c, k, keys = svo.extract_voxels(D)
    rc, rk = ref.extract_voxels(D)
    assert c.shape == rc.shape and c.shape[0] == pts.shape[0]
    assert float_bits_equal(c, rc)
    assert float_bits_equal(k, rk)


@pytest.mark.parametrize("res", [(96, 72), (160, 120)])
def test_raycast_matches_reference_bit_exact(P, res):
    """Same pool on both sides (uploaded), so the image comparison is independent of the integrate races."""
    w, h = res
    D = 7
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(96, 72)
    depth, rgb = P.synth.make_frame(96, 72, None, seed=1)
    svo = P.SVO(center, half, D)
    for _ in range(66):
        svo.integrate_depth(depth, rgb, fx, fy)
    ref = R.RefSVO(center, half, D)
    ref.load(svo.pool())
    rng = np.random.default_rng(3)
    for view in (LOOK_PLUS_Z, view_for_pose(random_pose(rng, 0.2, 0.3)), np.eye(4, dtype=np.float32)):
        img = svo.raycast(w, h, 45.0, view)
        want, _ = ref.raycast(w, h, 45.0, view)
        assert np.array_equal(img, want), "%d pixels differ" % np.count_nonzero(np.any(img != want, axis=2))
