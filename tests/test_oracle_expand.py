"""Map growth (osl_svo_expand / orc_svo_expand; reference octree.cpp:183-206, 362-378, quirk Q10).

The reference cannot expand a GPU-backed tree, so there is no reference output to pin against.  What pins the oracle's
restatement: a hand-derived known answer on the 3-point tree of SURVEY.md section 8c, and two size-independent
properties -- the occupied-voxel set is unchanged by an expansion, and equals that of a tree built from scratch in
the doubled cube one level deeper."""
import numpy as np

from oracle import oracle as orc
from common import EMPTY, FLAG, check_pool_invariants, pkg

PTS = np.array([[0.3, 0.3, 0.3], [0.9, 0.9, 0.9], [-0.9, 0.2, 0.6]], dtype=np.float32)
RGB = np.array([[200, 100, 50], [10, 20, 30], [255, 255, 255]], dtype=np.uint8)


def test_expand_known_answer():
    t = orc.OracleSVO((0, 0, 0), 1.0, 2)
    t.integrate_points(PTS, RGB)
    before = t.pool()
    assert t.size == 24
    t.expand(1)
    assert t.size == 88 and t.max_depth == 3 and t.half_edge == 2.0
    p = t.pool()
    w0, w1 = p[0::2], p[1::2]
    for i in range(8):
        tile = 24 + 8 * i
        assert w0[i] == FLAG | tile                    # new node i points at its tile
        moved = tile + 7 - i                           # old child i sits in the octant that touches the centre
        assert w0[moved] == before[2 * i] and w1[moved] == before[2 * i + 1]
        for k in range(8):
            if tile + k != moved:
                assert w0[tile + k] == 0 and w1[tile + k] == EMPTY
    # value of new node 7: children = {old node 7 (0x8105070D), 7 x empty}: channels >> 3, alpha = max
    assert w1[7] == 0x81000001                         # (0x0D >> 3, 0x07 >> 3, 0x05 >> 3, max(0x81, 0x7F))
    assert w1[6] == 0x81020202                         # old 0x81101010
    assert w1[1] == 0x7F000000                         # old node 1 was never touched (0): alpha = max(0, 127)
    # the tiles of the old tree are untouched
    assert np.array_equal(p[16:48], before[16:48])
    check_pool_invariants(p)


def test_expand_rejects_bad_layers():
    t = orc.OracleSVO((0, 0, 0), 1.0, 19)
    t.expand(1)
    try:
        t.expand(1)
    except ValueError:
        return
    raise AssertionError("depth 21 accepted")


def test_expand_empty_tree_only_changes_geometry():
    t = orc.OracleSVO((0, 0, 0), 1.0, 4)
    t.expand(2)
    assert t.size == 0 and t.max_depth == 6 and t.half_edge == 4.0
    t.integrate_points(PTS * 3.5, RGB)
    ref = orc.OracleSVO((0, 0, 0), 4.0, 6)
    ref.integrate_points(PTS * 3.5, RGB)
    assert np.array_equal(t.pool(), ref.pool())


def _frames(n, w=96, h=72):
    synth = pkg().synth
    fx, fy = synth.focal(w, h)
    out = []
    for k in range(n):
        pose = synth.orbit_pose(40 * k)
        d, c = synth.make_frame(w, h, pose, seed=k)
        out.append((d, c, fx, fy, pose))
    return out


def _voxel_set(t):
    cen, col, _ = t.extract_voxels()
    order = np.lexsort((cen[:, 2], cen[:, 1], cen[:, 0]))
    return cen[order], col[order]


def test_expand_keeps_the_voxel_set_and_matches_a_tree_built_in_the_larger_cube():
    frames = _frames(3)
    a = orc.OracleSVO((0, 0, 0), 4.0, 6)      # the room (|x|,|y|,|z| <= 2.5 m) fits: no clamped points
    b = orc.OracleSVO((0, 0, 0), 8.0, 7)
    for d, c, fx, fy, pose in frames[:2]:
        a.integrate_depth(d, c, fx, fy, pose)
        b.integrate_depth(d, c, fx, fy, pose)
    cen0, col0 = _voxel_set(a)
    n0 = a.size
    a.expand(1)
    assert a.size == n0 + 64
    check_pool_invariants(a.pool())
    cen1, col1 = _voxel_set(a)
    assert cen0.shape[0] > 500
    assert np.array_equal(cen0, cen1) and np.array_equal(col0, col1)
    cenb, colb = _voxel_set(b)
    assert np.array_equal(cen1, cenb) and np.array_equal(col1, colb)
    # and both keep agreeing when the map is extended after the expansion, outside the old cube too
    d, c, fx, fy, pose = frames[2]
    far = pose.copy()
    far[0, 3] += 4.0
    for t in (a, b):
        t.integrate_depth(d, c, fx, fy, pose)
        t.integrate_depth(d, c, fx, fy, far)
    cen1, col1 = _voxel_set(a)
    cenb, colb = _voxel_set(b)
    assert cen1[:, 0].max() > 4.0
    assert np.array_equal(cen1, cenb) and np.array_equal(col1, colb)


def test_expand_properties_on_random_clouds():
    """seeded random clouds, several depths and layer counts: growing keeps every occupied voxel (centre and colour),
    and growing then inserting equals inserting into a tree that was created large"""
    from common import unique_voxel_points
    for seed, (D, layers) in enumerate([(3, 1), (5, 2), (6, 3), (8, 1), (4, 2)]):
        rng = np.random.default_rng(100 + seed)
        half = 1.0
        a = orc.OracleSVO((0, 0, 0), half, D)
        b = orc.OracleSVO((0, 0, 0), half * 2 ** layers, D + layers)
        pts = unique_voxel_points(rng, 400, (0, 0, 0), half, D)
        rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
        for t in (a, b):
            t.integrate_points(pts, rgb)
            t.integrate_points(pts, rgb)          # alpha > 127 everywhere that is occupied
        before = _voxel_set(a)
        n0 = a.size
        a.expand(layers)
        assert a.size == n0 + 64 * layers and a.max_depth == D + layers and a.half_edge == half * 2 ** layers
        check_pool_invariants(a.pool())
        after = _voxel_set(a)
        assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
        big = _voxel_set(b)
        assert np.array_equal(after[0], big[0]) and np.array_equal(after[1], big[1])
        more = unique_voxel_points(rng, 300, (0, 0, 0), half * 2 ** layers, D + layers)
        rgb2 = rng.integers(0, 256, size=(more.shape[0], 3)).astype(np.uint8)
        for t in (a, b):
            t.integrate_points(more, rgb2)
            t.integrate_points(more, rgb2)
        ga, gb = _voxel_set(a), _voxel_set(b)
        assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1]), (seed, D, layers)
