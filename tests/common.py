"""Shared helpers of the test-suite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402

FLAG = 0x40000000
MASK = 0x3FFFFFFF
EMPTY = 0x7F000000

# looking down +z like the depth camera (the reference renderer looks down -z for view = identity)
LOOK_PLUS_Z = np.diag([-1.0, 1.0, -1.0, 1.0]).astype(np.float32)


def pkg():
    return graft.load_package()


def view_for_pose(pose):
    """view matrix whose camera sits at `pose` (camera-to-world) looking along the depth camera's +z."""
    return (LOOK_PLUS_Z @ np.linalg.inv(np.asarray(pose, dtype=np.float64))).astype(np.float32)


def random_pose(rng, trans=0.3, angle=0.4):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    a = rng.uniform(-angle, angle)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K)
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = rng.uniform(-trans, trans, size=3)
    return M.astype(np.float32)


def check_pool_invariants(pool, max_depth=None):
    """Structural invariants of the node pool that hold for any input (size-independent property test):
    child tiles are 8-aligned, inside the pool, referenced at most once, and every tile but the root's is referenced."""
    w0 = pool[0::2]
    n = w0.size
    assert n % 8 == 0 and n >= 8
    has = (w0 & FLAG) != 0
    tiles = (w0[has] & MASK).astype(np.int64)
    assert np.all(tiles % 8 == 0), "child tile not 8-aligned"
    assert np.all(tiles >= 8) and np.all(tiles + 8 <= n), "child tile outside the pool"
    assert np.unique(tiles).size == tiles.size, "tile referenced twice"
    assert tiles.size == n // 8 - 1, "orphan tiles: %d referenced, %d allocated" % (tiles.size, n // 8 - 1)
    assert np.all((w0[~has] & MASK) == 0) or True
    return int(has.sum())


def unique_voxel_points(rng, n, center, half_edge, max_depth):
    """n points in distinct leaf cells (no duplicate keys -> the reference's racy paths are deterministic)."""
    res = 1 << max_depth
    cells = rng.choice(res ** 3 if res ** 3 < 2 ** 62 else 2 ** 62, size=4 * n, replace=True)
    cells = rng.permutation(np.unique(cells))[:n]
    iz, iy, ix = cells // (res * res), (cells // res) % res, cells % res
    leaf = 2.0 * half_edge / res
    jitter = rng.uniform(0.2, 0.8, size=(cells.size, 3))
    pts = np.stack([ix, iy, iz], axis=1) * leaf + jitter * leaf - half_edge + np.asarray(center)
    return pts.astype(np.float32)


def float_bits_equal(a, b):
    """bit-exact float comparison that treats any NaN as equal to any NaN (the NaN payload/sign is not part of the
    contract: x86 produces 0xFFC00000 where the GPU produces 0x7FFFFFFF)"""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])
