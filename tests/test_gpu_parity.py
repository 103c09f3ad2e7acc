"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact: Morton keys, node indices, node words, raycast pixels (integer / byte work and float work whose operation
order is reproduced exactly; no tolerance is needed anywhere in this file)."""
import numpy as np
import pytest

from common import float_bits_equal, EMPTY, FLAG, LOOK_PLUS_Z, check_pool_invariants, pkg, random_pose, unique_voxel_points, view_for_pose
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    return pkg()


def test_library_loads_and_reports_version(P):
    assert b"sm_100a" in P.lib().osl_version()


# ------------------------------------------------------------------------------------------------ KAT
def test_kat_survey_8c(P):
    pts = np.array([[0.3, 0.3, 0.3], [0.9, 0.9, 0.9], [-0.9, 0.2, 0.6]], dtype=np.float32)
    rgb = np.array([[200, 100, 50], [10, 20, 30], [255, 255, 255]], dtype=np.uint8)
    keys = P.computeKeys(pts, (0, 0, 0), 1.0, 2)
    assert [oct(int(k)) for k in keys] == ["0o170", "0o177", "0o164"]
    t = P.SVO((0, 0, 0), 1.0, 2)
    t.integrate_points(pts, rgb)
    assert t.size == 24
    p = t.pool()
    w0, w1 = p[0::2], p[1::2]
    assert w0[6] == 0x40000008 and w0[7] == 0x40000010
    assert w1[12] == 0x81808080 and w1[16] == 0x81193264 and w1[23] == 0x810F0A05
    assert w1[6] == 0x81101010 and w1[7] == 0x8105070D
    assert w1[0] == 0x81020203
    t.integrate_points(pts, rgb)
    assert t.size == 32 and t.pool()[2 * 23] == 0x40000018  # Q3
    t2 = P.SVO((0, 0, 0), 1.0, 2, quirks=False)
    t2.integrate_points(pts, rgb)
    t2.integrate_points(pts, rgb)
    assert t2.size == 24


# ------------------------------------------------------------------------------------------------ keys / image kernels
@pytest.mark.parametrize("D", [1, 2, 8, 10, 11, 16, 20])
@pytest.mark.parametrize("stride", [3, 4])
def test_compute_keys(P, D, stride):
    rng = np.random.default_rng(100 + D)
    n = 20000
    pts = rng.uniform(-1.3, 1.3, size=(n, stride)).astype(np.float32)
    pts[::97, 0] = np.inf
    pts[1::97, 1] = np.nan  # Q1: y is not tested
    pts[2::97, 2] = -np.inf
    pts[3::97] = 0.0  # exactly on the centre planes (strict >)
    center, half = (0.1, -0.2, 0.05), 1.0
    got = P.computeKeys(pts, center, half, D)
    want = orc.compute_keys(pts, center, half, D)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("w,h", [(160, 120), (333, 77)])
def test_vertex_map_and_transform(P, w, h):
    rng = np.random.default_rng(5)
    depth = rng.integers(0, 16000, size=(h, w)).astype(np.uint16)
    depth[0, :7] = 0
    fx, fy = P.synth.focal(w, h)
    pose = random_pose(rng)
    pts = P.generateVertexMap(depth, fx, fy)
    want = orc.vertex_map(depth, fx, fy)
    got = pts.cpu().numpy()
    assert float_bits_equal(got, want)
    P.transformVertexMap(pts, pose)
    want_t = orc.transform(want, pose)
    assert float_bits_equal(pts.cpu().numpy(), want_t)
    box = P.computePointCloudBoundingBox(pts)
    assert np.array_equal(box, orc.bbox(want_t))


# ------------------------------------------------------------------------------------------------ integrate
def _run_frames(P, D, w, h, frames, quirks=True, reserve=0, same_frame=False):
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D, reserve_nodes=reserve, quirks=quirks)
    ref = orc.OracleSVO(center, half, D, quirks=quirks)
    for k in range(frames):
        kk = 0 if same_frame else k
        pose = P.synth.orbit_pose(25 * kk)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=kk)
        svo.integrate_depth(depth, rgb, fx, fy, pose)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
        assert svo.size == ref.size, "frame %d: size %d != %d" % (k, svo.size, ref.size)
        c, rc = svo.counters(), ref.counters()
        assert c.n_nodes == ref.size
    got, want = svo.pool(), ref.pool()
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, "first mismatching words %s" % bad[:10]
    return svo, ref


@pytest.mark.parametrize("D,w,h,frames", [(8, 160, 120, 4), (5, 64, 48, 3), (10, 320, 240, 3), (14, 160, 120, 3),
                                          (16, 320, 240, 2), (20, 96, 72, 2)])
def test_integrate_depth_matches_oracle(P, D, w, h, frames):
    svo, ref = _run_frames(P, D, w, h, frames)
    check_pool_invariants(svo.pool())
    c, rc = svo.counters(), ref.counters()
    # per-call counters on our side, running totals on the oracle's
    assert c.n_unique <= c.n_valid <= c.n_points == w * h


@pytest.mark.parametrize("w,h", [(333, 77), (70, 33), (8, 8), (64, 32), (65, 31), (1, 1), (640, 480)])
def test_integrate_depth_ragged_image_sizes(P, w, h):
    """tile edges of the 64x32-pixel emit tiles, widths that defeat the 128-bit depth loads"""
    _run_frames(P, 9, w, h, 2)


def test_pipelined_mode_matches_oracle(P):
    """emit + sort of frame f+1 on the front stream, overlapped with the tree update of frame f"""
    import torch
    D, w, h, frames = 12, 320, 240, 12
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D).set_pipeline(True)
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for k in range(frames):
        pose = P.synth.orbit_pose(7 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        d, c = torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()
        keep.append((d, c))
        ref.integrate_depth(depth, rgb, fx, fy, pose)
    torch.cuda.synchronize()  # the promise of pipelined mode: inputs are complete when the call is made
    for k in range(frames):
        svo.integrate_depth(keep[k][0], keep[k][1], fx, fy, P.synth.orbit_pose(7 * k))
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


def test_host_frames_pipeline_matches_oracle(P):
    D, w, h, frames = 10, 320, 240, 9
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D)
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for k in range(frames):
        pose = P.synth.orbit_pose(11 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        keep.append((depth, rgb))  # host buffers must stay untouched until the frame completes
        svo.integrate_depth_host(depth, rgb, fx, fy, pose)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


def test_bucket_sort_slow_path_after_a_scene_cut(P):
    """splitters come from an earlier frame: a frame whose keys all fall into ONE splitter range (> 2048 of them)
    takes k_sort_bucket's global-memory path; a frame of the old distribution afterwards takes the fast path again"""
    rng = np.random.default_rng(8)
    D = 9
    svo = P.SVO((0, 0, 0), 1.0, D)
    grid = P.SVO((0, 0, 0), 1.0, D, force_grid_sort=True)
    ref = orc.OracleSVO((0, 0, 0), 1.0, D)

    def cloud(n, lo, hi):
        pts = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
        return pts, rng.integers(0, 256, size=(n, 3)).astype(np.uint8)

    frames = [cloud(6000, -0.9, -0.1)] * 3 + [cloud(30000, 0.1, 0.9)] + [cloud(6000, -0.9, 0.9)] * 2
    for pts, rgb in frames:
        for t in (svo, grid, ref):
            t.integrate_points(pts, rgb)
        svo.sync()  # completed frames feed the grid-size / sort-selection hints
        assert svo.size == ref.size == grid.size
    assert np.array_equal(svo.pool(), ref.pool())
    assert np.array_equal(grid.pool(), ref.pool())


def test_stage_times_are_reported(P):
    D, w, h = 8, 160, 120
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    depth, rgb = P.synth.make_frame(w, h, None, seed=3)
    svo = P.SVO(center, half, D).set_stage_timing(True)
    svo.integrate_depth(depth, rgb, fx, fy)
    ms = svo.stage_times()
    assert len(ms) == 4 and all(0.0 < x < 100.0 for x in ms)


def test_duplicate_points_lowest_index_wins_across_tiles(P):
    """the same leaf hit from different emit tiles (inputs 2048 apart): canonical Q7 = lowest input index"""
    rng = np.random.default_rng(4)
    n = 3 * 2048 + 17
    base = rng.uniform(-0.9, 0.9, size=(37, 3)).astype(np.float32)
    pts = base[rng.integers(0, 37, size=n)]
    rgb = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
    svo = P.SVO((0, 0, 0), 1.0, 6)
    ref = orc.OracleSVO((0, 0, 0), 1.0, 6)
    svo.integrate_points(pts, rgb)
    ref.integrate_points(pts, rgb)
    assert np.array_equal(svo.pool(), ref.pool())
    assert svo.counters().n_unique == ref.counters().n_unique <= 37


def test_integrate_same_frame_repeatedly_q3_and_alpha(P):
    svo, ref = _run_frames(P, 8, 160, 120, 5, same_frame=True)
    w1 = svo.pool()[1::2]
    assert (w1 >> 24).max() == 127 + 2 * 5


def test_integrate_without_quirks(P):
    _run_frames(P, 8, 160, 120, 3, quirks=False, same_frame=True)


def test_pool_growth_from_tiny_reserve(P):
    svo, _ = _run_frames(P, 9, 160, 120, 3, reserve=16)
    assert svo.size > 16


def test_counters_match_oracle_single_frame(P):
    D, w, h = 8, 160, 120
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    depth, rgb = P.synth.make_frame(w, h, None, seed=3)
    svo = P.SVO(center, half, D)
    ref = orc.OracleSVO(center, half, D)
    svo.integrate_depth(depth, rgb, fx, fy)
    ref.integrate_depth(depth, rgb, fx, fy)
    c, rc = svo.counters(), ref.counters()
    assert (c.n_points, c.n_valid, c.n_unique, c.n_split) == (rc.n_points, rc.n_valid, rc.n_unique, rc.n_split)
    assert list(c.pass_sizes)[:D] == list(rc.pass_sizes)[:D]
    # oracle parents[] is indexed by mip pass (pass p handles depth D-1-p); ours by depth
    assert [c.parents[D - 1 - p] for p in range(D)] == list(rc.parents)[:D]
    assert c.algorithmic_bytes == 5 * w * h + 8 * c.n_unique + 68 * c.n_split + 68 * sum(list(c.parents)[:D])


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 2047, 2048, 2049, 5000])
def test_integrate_points_ragged_sizes(P, n):
    rng = np.random.default_rng(n)
    D = 6
    pts = rng.uniform(-1.1, 1.1, size=(n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
    if n > 3:
        pts[1] = np.inf
        pts[3] = pts[2]  # duplicate key
    svo = P.SVO((0, 0, 0), 1.0, D)
    ref = orc.OracleSVO((0, 0, 0), 1.0, D)
    for _ in range(2):
        svo.integrate_points(pts, rgb)
        ref.integrate_points(pts, rgb)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


def test_all_invalid_frame(P):
    D, w, h = 8, 64, 48
    center, half = P.synth.tree_params(D)
    depth = np.zeros((h, w), dtype=np.uint16)
    rgb = np.zeros((h, w, 3), dtype=np.uint8)
    svo = P.SVO(center, half, D)
    ref = orc.OracleSVO(center, half, D)
    svo.integrate_depth(depth, rgb, 100.0, 100.0)
    ref.integrate_depth(depth, rgb, 100.0, 100.0)
    assert svo.size == ref.size == 8
    assert np.array_equal(svo.pool(), ref.pool())


def test_out_of_cube_points_fold_into_boundary_cells(P):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-5, 5, size=(3000, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(3000, 3)).astype(np.uint8)
    svo = P.SVO((0, 0, 0), 1.0, 7)
    ref = orc.OracleSVO((0, 0, 0), 1.0, 7)
    svo.integrate_points(pts, rgb)
    ref.integrate_points(pts, rgb)
    assert np.array_equal(svo.pool(), ref.pool())


def test_alpha_saturates_at_255(P):
    pts = np.array([[0.3, 0.3, 0.3]], dtype=np.float32)
    rgb = np.array([[10, 200, 30]], dtype=np.uint8)
    svo = P.SVO((0, 0, 0), 1.0, 3)
    ref = orc.OracleSVO((0, 0, 0), 1.0, 3)
    for _ in range(70):
        svo.integrate_points(pts, rgb)
        ref.integrate_points(pts, rgb)
    assert np.array_equal(svo.pool(), ref.pool())
    assert (svo.pool()[1::2] >> 24).max() == 255


# ------------------------------------------------------------------------------------------------ voxel grid path
@pytest.mark.parametrize("D,n", [(6, 4000), (9, 20000)])
def test_integrate_voxels_matches_oracle(P, D, n):
    rng = np.random.default_rng(17)
    centers = np.ones((n, 4), dtype=np.float32)
    centers[:, :3] = rng.uniform(-0.95, 0.95, size=(n, 3))
    centers[::53, 0] = np.inf
    colors = rng.uniform(0, 1, size=(n, 4)).astype(np.float32)
    colors[::11] = 1.0  # Q15: 1.0 * 256 spills into the next channel
    svo = P.SVO((0, 0, 0), 1.0, D)
    ref = orc.OracleSVO((0, 0, 0), 1.0, D)
    for _ in range(2):
        svo.integrate_voxels(centers, colors)
        ref.integrate_voxels(centers, colors)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


# ------------------------------------------------------------------------------------------------ extraction
def test_extract_voxels_matches_oracle(P):
    svo, ref = _run_frames(P, 7, 160, 120, 2)
    for depth in (7, 5, 1):
        c, k, keys = svo.extract_voxels(depth)
        rc, rk, rkeys = ref.extract_voxels(depth)
        assert np.array_equal(keys, rkeys)
        assert np.all(np.diff(keys) > 0)  # sortedness
        assert float_bits_equal(c, rc)
        assert float_bits_equal(k, rk)


# ------------------------------------------------------------------------------------------------ raycast
def _saturated_tree(P, D=7, w=96, h=72, reps=66):
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    depth, rgb = P.synth.make_frame(w, h, None, seed=1)
    svo = P.SVO(center, half, D)
    for _ in range(reps):
        svo.integrate_depth(depth, rgb, fx, fy)
    return svo, center, half


@pytest.mark.parametrize("mode", [0, 1])
def test_raycast_matches_oracle(P, mode):
    svo, center, half = _saturated_tree(P)
    pool = svo.pool()
    rng = np.random.default_rng(3)
    for view in (LOOK_PLUS_Z, view_for_pose(random_pose(rng, 0.2, 0.3)), np.eye(4, dtype=np.float32)):
        st, cnt = P.RaycastStats(), orc.Counters()
        img = svo.raycast(96, 72, 45.0, view, mode=mode, stats=st)
        want = orc.raycast(pool, center, half, 96, 72, 45.0, view, mode=mode, counters=cnt)
        assert np.array_equal(img, want), "%d pixels differ" % np.count_nonzero(np.any(img != want, axis=2))
        assert (st.steps, st.visits) == (cnt.ray_steps, cnt.ray_visits)
    assert img.shape == (72, 96, 4)


def test_raycast_sees_the_scene(P):
    svo, center, half = _saturated_tree(P)
    img = svo.raycast(96, 72, 45.0, LOOK_PLUS_Z)
    assert np.all(img[..., 3] == 255)
    assert np.count_nonzero(img[..., :3].sum(axis=2)) > 0.5 * 96 * 72


def test_raycast_other_resolution_and_fov(P):
    svo, center, half = _saturated_tree(P)
    pool = svo.pool()
    img = svo.raycast(131, 57, 60.0, LOOK_PLUS_Z)
    want = orc.raycast(pool, center, half, 131, 57, 60.0, LOOK_PLUS_Z)
    assert np.array_equal(img, want)


def test_raycast_rows_tile_the_full_image(P):
    """the multi-GPU decomposition of the raycast: interleaved row bands rendered separately == the full image"""
    import torch
    svo, center, half = _saturated_tree(P)
    w, h = 96, 72
    full = svo.raycast(w, h, 45.0, LOOK_PLUS_Z)
    img = np.zeros_like(full)
    for world in (3,):
        for rank in range(world):
            for row0, rows in P.shard.row_bands(h, world, rank):
                out = torch.empty((rows, w, 4), dtype=torch.uint8, device="cuda")
                svo.raycast_rows(out, w, h, row0, rows, 45.0, LOOK_PLUS_Z)
                img[row0:row0 + rows] = out.cpu().numpy()
    assert np.array_equal(img, full)
    # the same decomposition with ONE launch per rank (osl_raycast_bands)
    for world in (1, 2, 5):
        img2 = np.zeros_like(full)
        band = P.shard.band_height(h, world)
        for rank in range(world):
            bands = P.shard.row_bands(h, world, rank, band)
            rows = sum(r for _, r in bands)
            out = torch.empty((max(rows, 1), w, 4), dtype=torch.uint8, device="cuda")
            assert svo.raycast_bands(out, w, h, band, world, rank, 45.0, LOOK_PLUS_Z) == rows
            got, off = out.cpu().numpy(), 0
            for row0, r in bands:
                img2[row0:row0 + r] = got[off:off + r]
                off += r
        assert np.array_equal(img2, full)


@pytest.mark.parametrize("D,res", [(12, (320, 240)), (16, (160, 120))])
def test_raycast_deep_tree_matches_oracle(P, D, res):
    """deep trees exercise the cached-ancestor descent (most steps resume 3-4 levels above the sample)"""
    w, h = res
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D)
    for k in range(3):
        pose = P.synth.orbit_pose(40 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        for _ in range(22):
            svo.integrate_depth(depth, rgb, fx, fy, pose)
    pool = svo.pool()
    for k in (0, 1):
        view = view_for_pose(P.synth.orbit_pose(40 * k))
        for mode in (0, 1):
            st, cnt = P.RaycastStats(), orc.Counters()
            img = svo.raycast(w, h, 45.0, view, mode=mode, stats=st)
            want = orc.raycast(pool, center, half, w, h, 45.0, view, mode=mode, counters=cnt)
            assert np.array_equal(img, want), "%d pixels differ" % np.count_nonzero(np.any(img != want, axis=2))
            assert (st.steps, st.visits) == (cnt.ray_steps, cnt.ray_visits)


def test_bench_workload_pipelined_matches_oracle(P):
    """BASELINE's headline configuration (640x480 orbit into a depth-16 SVO) at full frame size, pipelined with host
    frames AND with resident frames, against the oracle on the first frames of the orbit; then size-independent
    properties of the pool"""
    import torch
    D, w, h, frames = 16, 640, 480, 6
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    host = P.SVO(center, half, D, reserve_nodes=1 << 22)
    dev = P.SVO(center, half, D, reserve_nodes=16).set_pipeline(True)  # tiny reserve: the pool grows while piped
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for k in range(frames):
        pose = P.synth.orbit_pose(k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        keep.append((depth, rgb, torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()))
        ref.integrate_depth(depth, rgb, fx, fy, pose)
    torch.cuda.synchronize()
    for k in range(frames):
        pose = P.synth.orbit_pose(k)
        host.integrate_depth_host(keep[k][0], keep[k][1], fx, fy, pose)
        dev.integrate_depth(keep[k][2], keep[k][3], fx, fy, pose)
    want = ref.pool()
    for svo in (host, dev):
        assert svo.size == ref.size
        got = svo.pool()
        assert np.array_equal(got, want)
        check_pool_invariants(got)
    # raycast of the fused map, full resolution, bit-exact against the oracle on a row band (the CPU oracle marches
    # ~1 Mray/s): rows 200..215
    view = view_for_pose(P.synth.orbit_pose(frames - 1))
    out = torch.empty((16, w, 4), dtype=torch.uint8, device="cuda")
    dev.raycast_rows(out, w, h, 200, 16, 45.0, view)
    full = dev.raycast(w, h, 45.0, view)
    assert np.array_equal(out.cpu().numpy(), full[200:216])
    # every voxel the extraction reports is a leaf with alpha > 127 and vice versa (count check)
    c, k, keys = dev.extract_voxels(D)
    assert np.all(np.diff(keys) > 0)


def test_long_pipelined_sequence_properties(P):
    """200 pipelined frames of the bench workload (no oracle: size-independent properties only): the frame counter,
    monotone node count, structural invariants, alpha saturation bound, idempotent structure on repeated frames"""
    import torch
    D, w, h = 16, 640, 480
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    ring = []
    for k in range(8):
        pose = P.synth.orbit_pose(3 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        ring.append((torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), pose))
    torch.cuda.synchronize()
    svo = P.SVO(center, half, D, reserve_nodes=1 << 22).set_pipeline(True)
    sizes = []
    for k in range(200):
        d, c, pose = ring[k % 8]
        svo.integrate_depth(d, c, fx, fy, pose)
        if k % 40 == 39:
            sizes.append(svo.size)
    cn = svo.counters()
    assert cn.frames == 200
    assert all(a <= b for a, b in zip(sizes, sizes[1:]))
    # the 8 frames repeat: after the first laps (Q3 splits a leaf once, on its second observation) the structure is
    # a fixed point -- no further nodes
    assert sizes[-1] == sizes[-2]
    pool = svo.pool()
    check_pool_invariants(pool)
    alpha = pool[1::2] >> 24
    assert alpha.max() <= 255 and alpha.max() >= 127 + 2 * 25  # every leaf of a frame is seen 25 times


def test_cfg4_frame_size_1280x960_depth16_matches_oracle(P):
    """BASELINE configs[3] frame size: 1280x960 into a depth-16 SVO (2 frames; the oracle needs ~1 s per frame)"""
    D, w, h = 16, 1280, 960
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D, reserve_nodes=1 << 22)
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for k in range(2):
        pose = P.synth.orbit_pose(5 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        keep.append((depth, rgb))
        svo.integrate_depth_host(depth, rgb, fx, fy, pose)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())
    c, rc = svo.counters(), ref.counters()
    assert c.n_points == w * h


def test_cfg3_orbit_depth14_incremental_matches_oracle(P):
    """BASELINE configs[2] shape (incremental fusion along the orbit, depth-14 SVO), first 12 frames at 320x240
    against the oracle, pipelined"""
    import torch
    D, w, h, frames = 14, 320, 240, 12
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D).set_pipeline(True)
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for k in range(frames):
        pose = P.synth.orbit_pose(k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        keep.append((torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()))
        ref.integrate_depth(depth, rgb, fx, fy, pose)
    torch.cuda.synchronize()
    for k in range(frames):
        svo.integrate_depth(keep[k][0], keep[k][1], fx, fy, P.synth.orbit_pose(k))
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


def test_checkpoint_resume_continues_bit_exactly(P, tmp_path):
    """save after 3 frames, restore into a fresh tree, integrate 3 more on both: identical pools (and == oracle)"""
    D, w, h = 10, 160, 120
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    frames = []
    for k in range(6):
        pose = P.synth.orbit_pose(12 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        frames.append((depth, rgb, pose))
    a = P.SVO(center, half, D)
    ref = orc.OracleSVO(center, half, D)
    for depth, rgb, pose in frames[:3]:
        a.integrate_depth(depth, rgb, fx, fy, pose)
    path = tmp_path / "map.oslsvo"
    a.save(path)
    b = P.SVO(center, half, D)
    b.restore(path)
    assert b.size == a.size and np.array_equal(a.pool(), b.pool())
    for depth, rgb, pose in frames[3:]:
        a.integrate_depth(depth, rgb, fx, fy, pose)
        b.integrate_depth_host(depth, rgb, fx, fy, pose)
    for depth, rgb, pose in frames:
        ref.integrate_depth(depth, rgb, fx, fy, pose)
    assert np.array_equal(a.pool(), b.pool())
    assert np.array_equal(a.pool(), ref.pool())
    # a checkpoint of another tree geometry is refused
    c = P.SVO(center, half, D + 1)
    with pytest.raises(P.OslError):
        c.restore(path)


@pytest.mark.parametrize("variant", ["sorted", "sorted_with_duplicates", "sorted_with_invalid", "reversed"])
def test_voxel_grids_in_morton_order_skip_the_sort(P, variant):
    """a VoxelGrid that arrives sorted and gap-free is not sorted again (k_emit keeps a dense copy); every variant
    must still equal the oracle, which always sorts (svo.cu:602)"""
    rng = np.random.default_rng(5)
    D, n = 7, 6000
    pts = unique_voxel_points(rng, n, (0, 0, 0), 1.0, D)
    keys = orc.compute_keys(pts, (0, 0, 0), 1.0, D)
    pts = pts[np.argsort(keys, kind="stable")]
    if variant == "sorted_with_duplicates":
        pts[100:104] = pts[100]
        pts[2000] = pts[1999]
    elif variant == "sorted_with_invalid":
        pts[777, 0] = np.inf
    elif variant == "reversed":
        pts = pts[::-1].copy()
    centers = np.ones((pts.shape[0], 4), dtype=np.float32)
    centers[:, :3] = pts
    colors = rng.uniform(0, 1, size=centers.shape).astype(np.float32)
    svo = P.SVO((0, 0, 0), 1.0, D)
    ref = orc.OracleSVO((0, 0, 0), 1.0, D)
    for _ in range(2):
        svo.integrate_voxels(centers, colors)
        ref.integrate_voxels(centers, colors)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


def test_two_pipelined_trees_interleaved(P):
    """two maps fed in lock-step from one host thread, both pipelined (cooperative-grid budget is shared)"""
    import torch
    D, w, h, frames = 12, 320, 240, 10
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    a = P.SVO(center, half, D).set_pipeline(True)
    b = P.SVO(center, half, D)
    ra, rb = orc.OracleSVO(center, half, D), orc.OracleSVO(center, half, D)
    keep = []
    for k in range(frames):
        pa, pb = P.synth.orbit_pose(5 * k), P.synth.orbit_pose(300 + 5 * k)
        da, ca = P.synth.make_frame(w, h, pa, seed=k)
        db, cb = P.synth.make_frame(w, h, pb, seed=100 + k)
        keep.append((torch.from_numpy(da).cuda(), torch.from_numpy(ca).cuda(), db, cb))
        ra.integrate_depth(da, ca, fx, fy, pa)
        rb.integrate_depth(db, cb, fx, fy, pb)
    torch.cuda.synchronize()
    for k in range(frames):
        a.integrate_depth(keep[k][0], keep[k][1], fx, fy, P.synth.orbit_pose(5 * k))
        b.integrate_depth_host(keep[k][2], keep[k][3], fx, fy, P.synth.orbit_pose(300 + 5 * k))
    assert np.array_equal(a.pool(), ra.pool())
    assert np.array_equal(b.pool(), rb.pool())


@pytest.mark.parametrize("n,D", [(400000, 10), (700000, 13)])
def test_large_point_clouds_shared_walks(P, n, D):
    """enough unique keys that every k_structure CTA owns several 512-key blocks (contiguous ranges, shared tree
    walks, multi-tile radix sort), two frames so that the second one walks a real tree"""
    rng = np.random.default_rng(n)
    svo = P.SVO((0, 0, 0), 1.0, D, reserve_nodes=1 << 22)
    ref = orc.OracleSVO((0, 0, 0), 1.0, D)
    for frame in range(2):
        pts = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
        pts[: n // 3] *= 0.2  # a dense cluster: long shared prefixes next to sparse space
        rgb = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
        svo.integrate_points(pts, rgb)
        ref.integrate_points(pts, rgb)
        assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fuzz_mixed_modes_sizes_and_entry_points(P, seed):
    """random sequence of depth frames (device strict / device pipelined / host), point clouds and voxel grids of
    random sizes into ONE tree, with mode switches, syncs and raycasts in between -- against the oracle"""
    import torch
    rng = np.random.default_rng(100 + seed)
    D = int(rng.integers(6, 13))
    center, half = P.synth.tree_params(D)
    svo = P.SVO(center, half, D, reserve_nodes=int(rng.choice([16, 1 << 16])))
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for step in range(24):
        kind = rng.choice(["strict", "piped", "host", "points", "voxels"], p=[0.25, 0.25, 0.25, 0.15, 0.10])
        if kind in ("strict", "piped", "host"):
            w, h = [(64, 48), (160, 120), (200, 150), (320, 240)][int(rng.integers(0, 4))]
            fx, fy = P.synth.focal(w, h)
            pose = P.synth.orbit_pose(int(rng.integers(0, 400)))
            depth, rgb = P.synth.make_frame(w, h, pose, seed=int(rng.integers(0, 1000)))
            ref.integrate_depth(depth, rgb, fx, fy, pose)
            if kind == "host":
                keep.append((depth, rgb))
                svo.integrate_depth_host(depth, rgb, fx, fy, pose)
            else:
                d, c = torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()
                torch.cuda.synchronize()
                keep.append((d, c))
                svo.set_pipeline(kind == "piped")
                svo.integrate_depth(d, c, fx, fy, pose)
        elif kind == "points":
            n = int(rng.integers(0, 5000))
            pts = rng.uniform(-1.5 * half / 100, 1.5 * half / 100, size=(n, 3)).astype(np.float32)
            rgb = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
            svo.set_pipeline(bool(rng.integers(0, 2)))
            svo.integrate_points(pts, rgb)
            ref.integrate_points(pts, rgb)
        else:
            n = int(rng.integers(1, 3000))
            cen = np.ones((n, 4), dtype=np.float32)
            cen[:, :3] = rng.uniform(-half / 50, half / 50, size=(n, 3))
            col = rng.uniform(0, 1, size=(n, 4)).astype(np.float32)
            svo.set_pipeline(False)
            svo.integrate_voxels(cen, col)
            ref.integrate_voxels(cen, col)
        r = rng.random()
        if r < 0.2:
            assert svo.size == ref.size, "step %d (%s)" % (step, kind)
        elif r < 0.3:
            svo.raycast(32, 24, 45.0, LOOK_PLUS_Z)
        elif r < 0.35:
            svo.sync()
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


# ------------------------------------------------------------------------------------------------ map growth
@pytest.mark.parametrize("mode", ["strict", "pipelined", "host"])
def test_expand_then_keep_integrating_matches_oracle(P, mode):
    """osl_svo_expand (octree.cpp:183-206, 362-378 made to work, Q10): pool bit-exact against the oracle's restatement
    right after the expansion and after further frames, some of them outside the old cube; voxel set unchanged."""
    import torch
    D, w, h = 8, 160, 120
    center, half = (0.0, 0.0, 0.0), 4.0
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D).set_pipeline(mode == "pipelined")
    ref = orc.OracleSVO(center, half, D)
    keep = []

    def feed(k, shift=0.0):
        pose = P.synth.orbit_pose(23 * k)
        pose[0, 3] += shift
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
        if mode == "host":
            keep.append((depth, rgb))
            svo.integrate_depth_host(depth, rgb, fx, fy, pose)
        else:
            d, c = torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()
            keep.append((d, c))
            torch.cuda.synchronize()
            svo.integrate_depth(d, c, fx, fy, pose)

    for k in range(3):
        feed(k)
    cen0, col0, _ = svo.extract_voxels()
    svo.expand(1)
    ref.expand(1)
    assert (svo.max_depth, svo.half_edge) == (D + 1, 8.0) == (ref.max_depth, ref.half_edge)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())
    cen1, col1, keys1 = svo.extract_voxels()
    o0 = np.lexsort((cen0[:, 2], cen0[:, 1], cen0[:, 0]))
    o1 = np.lexsort((cen1[:, 2], cen1[:, 1], cen1[:, 0]))
    assert np.array_equal(cen0[o0], cen1[o1]) and np.array_equal(col0[o0], col1[o1])
    rc, rk, rkeys = ref.extract_voxels()
    assert np.array_equal(keys1, rkeys) and float_bits_equal(cen1, rc) and float_bits_equal(col1, rk)
    for k in range(3, 8):
        feed(k, 4.5 if k % 2 else 0.0)
    assert svo.size == ref.size
    pool = svo.pool()
    assert np.array_equal(pool, ref.pool())
    check_pool_invariants(pool)
    view = view_for_pose(P.synth.orbit_pose(0))
    img = svo.raycast(96, 72, 45.0, view, mode=1)
    assert np.array_equal(img, orc.raycast(pool, center, 8.0, 96, 72, 45.0, view, mode=1))
    # two more layers at once, then a frame
    svo.expand(2)
    ref.expand(2)
    feed(9, -9.0)
    assert (svo.max_depth, svo.half_edge) == (D + 3, 32.0)
    assert np.array_equal(svo.pool(), ref.pool())


def test_expand_empty_tree_and_depth_limit(P):
    svo = P.SVO((0, 0, 0), 1.0, 18)
    svo.expand(2)
    assert svo.size == 0 and svo.max_depth == 20 and svo.half_edge == 4.0
    with pytest.raises(Exception):
        svo.expand(1)
    pts = np.array([[0.3, 0.3, 0.3], [3.9, 3.9, 3.9], [-3.9, 0.2, 0.6]], dtype=np.float32)
    rgb = np.array([[200, 100, 50], [10, 20, 30], [255, 255, 255]], dtype=np.uint8)
    svo.integrate_points(pts, rgb)
    ref = orc.OracleSVO((0, 0, 0), 4.0, 20)
    ref.integrate_points(pts, rgb)
    assert np.array_equal(svo.pool(), ref.pool())


def test_octree_expand_by_size(P):
    """world::Octree::expandBySize (octree.cpp:362-378): enough doublings for size_ + add_size, resolution kept"""
    w, h = 96, 72
    fx, fy = P.synth.focal(w, h)
    depth, rgb = P.synth.make_frame(w, h, None, seed=4)
    oc = P.Octree(0.0625, (0.0, 0.0, 0.0), 4.0)
    oc.addDepthFrame(depth, rgb, fx, fy)
    D = oc.svo.max_depth
    n = oc.svo.size
    oc.expandBySize(0.0)
    assert oc.svo.size == n
    oc.expandBySize(5.0)  # 9 / 4 -> 2 doublings
    assert oc.size_ == 16.0 and oc.svo.max_depth == D + 2 and oc.svo.size == n + 128
    oc.addDepthFrame(depth, rgb, fx, fy)  # _max_depth(resolution) follows size_
    ref = orc.OracleSVO((0, 0, 0), 4.0, D)
    ref.integrate_depth(depth, rgb, fx, fy)
    ref.expand(2)
    ref.integrate_depth(depth, rgb, fx, fy)
    assert np.array_equal(oc.svo.pool(), ref.pool())


def test_checkpoint_of_an_expanded_tree(P, tmp_path):
    svo, ref = _run_frames(P, 7, 96, 72, 2)
    svo.expand(1)
    ref.expand(1)
    path = str(tmp_path / "grown.svo")
    svo.save(path)
    again = P.SVO(svo.view()[2], svo.half_edge, svo.max_depth)
    again.restore(path)
    assert again.size == ref.size and np.array_equal(again.pool(), ref.pool())
    with pytest.raises(Exception):
        P.SVO(svo.view()[2], svo.half_edge / 2, svo.max_depth - 1).restore(path)  # the geometry before the expansion


@pytest.mark.parametrize("seed,D,layers", [(0, 3, 1), (1, 5, 2), (2, 6, 3), (3, 9, 1), (4, 17, 3)])
def test_expand_random_clouds_match_oracle(P, seed, D, layers):
    rng = np.random.default_rng(200 + seed)
    half = 1.0
    svo, ref = P.SVO((0, 0, 0), half, D), orc.OracleSVO((0, 0, 0), half, D)
    pts = rng.uniform(-1.0, 1.0, size=(3000, 3)).astype(np.float32)     # duplicates included
    rgb = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
    for t in (svo, ref):
        t.integrate_points(pts, rgb)
        t.expand(layers)
    assert np.array_equal(svo.pool(), ref.pool())
    big = half * 2 ** layers
    more = rng.uniform(-big, big, size=(5000, 3)).astype(np.float32)
    rgb2 = rng.integers(0, 256, size=(more.shape[0], 3)).astype(np.uint8)
    for t in (svo, ref):
        t.integrate_points(more, rgb2)
        t.integrate_points(pts, rgb)
    assert svo.size == ref.size
    pool = svo.pool()
    assert np.array_equal(pool, ref.pool())
    check_pool_invariants(pool)
    c, k, keys = svo.extract_voxels()
    rc, rk, rkeys = ref.extract_voxels()
    assert np.array_equal(keys, rkeys) and float_bits_equal(c, rc) and float_bits_equal(k, rk)
