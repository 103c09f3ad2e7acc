"""GPU: inputs of MILLIONS of keys (the cfg2 / cfg5 size class) against the CPU oracle, bit for bit.  From 2^20 expected
keys on, svoFromVoxelGrid / svoFromPointCloud take the big-input variants of the kernels (k_structure_big: streamed
phase A in which only the keys that open a new parent walk the tree, prefetched phase C; k_levels: leaf-balanced subtree shares; the grid radix sort),
which the frame-sized tests never reach."""
import numpy as np
import pytest

from common import pkg
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    return pkg()


@pytest.fixture(scope="module")
def sphere_grid(P):
    """a sphere voxelised at depth 10: ~1.9 M voxels in Morton order (device tensors)"""
    D = 10
    V, T = P.synth.icosphere(5, 0.8, (0.05, -0.02, 0.03))
    colors = np.random.default_rng(3).uniform(0.0, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    cen, col = P.meshToVoxelGrid(V, T, colors, (0.0, 0.0, 0.0), 1.0, D)
    assert cen.shape[0] > (1 << 20)
    return D, cen, col


@pytest.mark.parametrize("order", ["morton", "shuffled", "morton_with_duplicates_and_invalid"])
def test_big_voxel_grid_matches_oracle(P, sphere_grid, order):
    import torch
    D, cen, col = sphere_grid
    n = cen.shape[0]
    if order == "shuffled":
        perm = torch.from_numpy(np.random.default_rng(9).permutation(n)).cuda()
        cen, col = cen[perm].contiguous(), col[perm].contiguous()
    elif order == "morton_with_duplicates_and_invalid":
        cen, col = cen.clone(), col.clone()
        cen[5000:5003] = cen[5000]          # runs of equal keys (Q11: colours go by sorted position)
        cen[n // 2] = cen[n // 2 - 1]
        cen[n - 1] = cen[n - 2]
        cen[123457, 1] = float("inf")       # an invalid voxel: the grid is no longer gap-free, the sort runs
    svo = P.SVO((0, 0, 0), 1.0, D, reserve_nodes=int(3.0 * n))
    ref = orc.OracleSVO((0, 0, 0), 1.0, D)
    cen_h, col_h = cen.cpu().numpy(), col.cpu().numpy()
    for k in range(3):  # build, re-observe with Q3 splits, steady state
        svo.integrate_voxels(cen, col)
        ref.integrate_voxels(cen_h, col_h)
        assert svo.size == ref.size, "observation %d" % k
    a, b = svo.pool(), ref.pool()
    assert np.array_equal(a[0::2], b[0::2]), "child pointers differ"
    assert np.array_equal(a, b), "%d values differ" % np.count_nonzero(a[1::2] != b[1::2])
    # a second, overlapping grid (shifted by a few cells): new sub-trees next to existing ones
    cen2 = cen.clone()
    cen2[:, 0] += 7.0 * 2.0 / (1 << D)
    svo.integrate_voxels(cen2, col)
    ref.integrate_voxels(cen2.cpu().numpy(), col_h)
    assert np.array_equal(svo.pool(), ref.pool())


def test_big_point_cloud_matches_oracle(P):
    """2.5 M points with many duplicates per leaf (depth 9): the de-duplicating emit, the grid sort on (key, index)
    pairs and the canonical lowest-index rule at this size"""
    rng = np.random.default_rng(11)
    D, n = 9, 2_500_000
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (0.75 * d + rng.normal(scale=0.002, size=(n, 3))).astype(np.float32)
    pts[::100003, 2] = np.nan
    rgb = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
    svo = P.SVO((0, 0, 0), 1.0, D, reserve_nodes=1 << 23)
    ref = orc.OracleSVO((0, 0, 0), 1.0, D)
    for k in range(3):
        svo.integrate_points(pts, rgb)
        ref.integrate_points(pts, rgb)
        assert svo.size == ref.size, "observation %d" % k
    assert svo.counters().n_unique > (1 << 19)
    assert np.array_equal(svo.pool(), ref.pool())


def test_big_voxel_grid_pipelined_matches_strict(P, sphere_grid):
    """the big-input kernels under the frame pipeline (internal streams, cooperative grids capped so that the stages of
    consecutive calls can be co-resident): same pool as the strict, stream-ordered calls"""
    import torch
    D, cen, col = sphere_grid
    n = cen.shape[0]
    perm = torch.from_numpy(np.random.default_rng(10).permutation(n)).cuda()
    cen_s, col_s = cen[perm].contiguous(), col[perm].contiguous()
    torch.cuda.synchronize()
    strict = P.SVO((0, 0, 0), 1.0, D, reserve_nodes=int(3.0 * n))
    piped = P.SVO((0, 0, 0), 1.0, D, reserve_nodes=int(3.0 * n)).set_pipeline(True)
    for a, b in ((cen, col), (cen_s, col_s), (cen, col), (cen_s, col_s)):
        strict.integrate_voxels(a, b)
        piped.integrate_voxels(a, b)
    piped.sync()
    assert strict.size == piped.size
    assert np.array_equal(strict.pool(), piped.pool())


@pytest.mark.parametrize("center,half,D", [((0.0, 0.0, 0.0), 1.0, 12), ((0.0123, -0.0456, 0.789), 1.3493741, 12),
                                           ((3.3, -7.7, 12.1), 0.7, 9), ((0.0, 0.0, 0.0), 655.36, 16),
                                           ((1000.5, 2000.25, -3000.125), 10.0, 20), ((0.1, 0.2, 0.3), 5.0, 3)])
def test_voxel_keys_on_cell_boundaries(P, center, half, D):
    """k_emit_grid decides most keys in closed form and certifies them; coordinates within a few ulp of a cell boundary go
    to the reference's float descent instead.  Adversarial inputs: coordinates exactly ON boundaries of every level, one
    ulp to either side, a few ulp away, NaN / INF in y (Q1: still a valid key), INF in x (invalid), points outside the
    cube -- through svoFromVoxelGrid on the GPU and in the oracle (which only knows the descent): equal pools."""
    rng = np.random.default_rng(D * 7 + 1)
    n = 120_000
    c = np.asarray(center, dtype=np.float64)
    lo = c - half
    pts = np.empty((n, 3), dtype=np.float32)
    for a in range(3):
        lvl = rng.integers(1, D + 1, size=n)
        k = (rng.random(n) * (2.0 ** lvl)).astype(np.int64)
        b = (lo[a] + k * (2.0 * half) / (2.0 ** lvl)).astype(np.float32)      # a cell boundary of level lvl, as a float
        kind = rng.integers(0, 8, size=n)
        up = np.nextafter(b, np.float32(np.inf))
        dn = np.nextafter(b, np.float32(-np.inf))
        far = b
        for _ in range(5):
            far = np.nextafter(far, np.float32(np.inf))
        rnd = (lo[a] + rng.random(n) * 2.0 * half).astype(np.float32)
        outside = (lo[a] + (rng.random(n) * 4.0 - 1.0) * 2.0 * half).astype(np.float32)
        pts[:, a] = np.select([kind == 0, kind == 1, kind == 2, kind == 3, kind == 4], [b, up, dn, far, outside], rnd)
    pts[::997, 1] = np.nan
    pts[5::997, 1] = np.inf
    pts[7::997, 1] = -np.inf
    pts[11::997, 0] = np.inf       # invalid voxels
    pts[13::997, 2] = np.nan
    centers = np.ones((n, 4), dtype=np.float32)
    centers[:, :3] = pts
    colors = rng.uniform(0, 1, size=(n, 4)).astype(np.float32)
    svo = P.SVO(center, half, D, reserve_nodes=1 << 22)
    ref = orc.OracleSVO(center, half, D)
    for _ in range(2):
        svo.integrate_voxels(centers, colors)
        ref.integrate_voxels(centers, colors)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())
    # the same points pre-sorted by key (no invalid ones): the "already in Morton order" decision must survive the fix-ups
    good = np.isfinite(pts[:, 0]) & np.isfinite(pts[:, 2])
    keys = orc.compute_keys(pts[good], center, half, D)
    order = np.argsort(keys, kind="stable")
    cs, ks = centers[good][order], colors[good][order]
    svo2 = P.SVO(center, half, D, reserve_nodes=1 << 22)
    ref2 = orc.OracleSVO(center, half, D)
    svo2.integrate_voxels(cs, ks)
    ref2.integrate_voxels(cs, ks)
    assert np.array_equal(svo2.pool(), ref2.pool())
