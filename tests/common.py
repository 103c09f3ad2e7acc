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


# ---- per-leaf legal-outcome check of one integrated frame (reference quirk Q7, svo.cu:366-381) -------------------

def blend_u8(cur, rgb):
    """fillNodes(Color256) (svo.cu:366-381) in integers: cur uint32 word1 [n], rgb uint8 [n, 3] -> new word1 [n]"""
    cur = np.asarray(cur, dtype=np.uint64)
    a = cur >> 24
    out = np.zeros_like(cur)
    for ch in range(3):
        c = (cur >> (8 * ch)) & 0xFF
        out |= ((((256 - a) * rgb[:, ch].astype(np.uint64) + a * c) >> 8) & 0xFF) << (8 * ch)
    return (out | (np.minimum(255, a + 2) << 24)).astype(np.uint32)


def average8(vals):
    """averageChildren with Q5 (svo.cu:384-441): vals uint32 [n, 8] -> uint32 [n]"""
    v = np.asarray(vals, dtype=np.uint64)
    out = np.zeros(v.shape[0], dtype=np.uint64)
    for ch in range(3):
        out |= (((v >> (8 * ch)) & 0xFF).sum(axis=1) >> 3) << (8 * ch)
    return (out | ((v >> 24).max(axis=1) << 24)).astype(np.uint32)


def descend(pool, keys, D):
    """node index at every depth 1..D for leading-1 Morton keys (int64 [n]) whose whole path exists in `pool`.
    Returns int64 [D, n] (row d-1 = the depth-d node)."""
    keys = np.asarray(keys, dtype=np.int64)
    w0 = pool[0::2]
    path = np.zeros((D, keys.size), dtype=np.int64)
    node = (keys >> (3 * (D - 1))) & 7
    path[0] = node
    for d in range(2, D + 1):
        w = w0[node]
        assert np.all(w & FLAG), "path of a key is not in the tree at depth %d" % (d - 1)
        node = (w & MASK).astype(np.int64) + ((keys >> (3 * (D - d))) & 7)
        path[d - 1] = node
    return path


def check_frame_outcome(before, after, keys, rgb, D, canonical):
    """One integrated depth frame, every leaf and every touched node checked.

    before / after: uint32 pools (2 words per node) around the frame; keys: int64 leading-1 key per input (1 =
    invalid); rgb: uint8 [n, 3].  canonical=True (ours): a leaf's value is the blend of the LOWEST input index that
    maps to it, alpha += 2 once.  canonical=False (the reference's racing read-modify-writes): the value must be
    reachable by m >= 1 successive blends of colours of inputs that map to the leaf, alpha += 2m.  Both: untouched
    value words are unchanged, every touched inner node holds averageChildren of its tile (node 0 apart: Q6)."""
    n_after, n_before = after.size // 2, before.size // 2
    bw1 = np.full(n_after, EMPTY, dtype=np.uint32)
    bw1[:n_before] = before[1::2]
    if n_before == 0:
        bw1[:8] = 0  # initOctree (svo.cu:24-31): the root's children start as {0, 0}
    aw1 = after[1::2]
    keys = np.asarray(keys, dtype=np.int64).ravel()
    rgb = np.asarray(rgb, dtype=np.uint8).reshape(-1, 3)
    valid = np.flatnonzero(keys != 1)
    order = valid[np.lexsort((valid, keys[valid]))]      # by key, then by input index
    ks = keys[order]
    head = np.ones(ks.size, dtype=bool)
    head[1:] = ks[1:] != ks[:-1]
    starts = np.flatnonzero(head)
    ukeys = ks[starts]
    counts = np.diff(np.append(starts, ks.size))
    path = descend(after, ukeys, D)
    leaf = path[D - 1]
    group = np.cumsum(head) - 1                           # leaf group of every sorted input
    cur = bw1[leaf]
    got = aw1[leaf]
    ca, ga = (cur >> 24).astype(int), (got >> 24).astype(int)
    if canonical:
        want = blend_u8(cur, rgb[order[starts]])
        bad = np.flatnonzero(want != got)
        assert bad.size == 0, "%d leaves are not the blend of their lowest input" % bad.size
    else:
        # number of read-modify-writes that landed on the leaf: exact while alpha has not saturated at 255
        sat = ga == 255
        steps = (ga - ca) // 2
        assert np.all((ga - ca)[~sat] % 2 == 0) and np.all(steps[~sat] >= 1) and np.all(steps <= counts), \
            "alpha outside {+2 .. +2k}"
        assert sat.mean() < 0.02, "too many saturated leaves for a meaningful check"
        one = blend_u8(cur[group], rgb[order]) == got[group]
        ok = np.zeros(ukeys.size, dtype=bool)
        np.logical_or.at(ok, group, one)
        ok &= (steps == 1)
        skipped = 0
        for g in np.flatnonzero(~ok & ~sat):                  # serialised duplicates: chains of `steps` blends
            cols = np.unique(rgb[order[starts[g]:starts[g] + counts[g]]], axis=0)
            states = np.array([cur[g]], dtype=np.uint32)
            for _ in range(int(steps[g])):
                if states.size * cols.shape[0] > 2_000_000:
                    states = None
                    break
                states = np.unique(blend_u8(np.repeat(states, cols.shape[0]), np.tile(cols, (states.size, 1))))
            if states is None:
                skipped += 1
                continue
            assert got[g] in states, "leaf %d: value %08x is no chain of %d blends of its %d inputs" % (
                leaf[g], got[g], steps[g], counts[g])
        assert skipped <= 0.01 * ukeys.size, "%d leaves too expensive to verify" % skipped
    # inner nodes on the touched paths: averageChildren of their tile; node 0's value word is Q6's
    touched = np.zeros(n_after, dtype=bool)
    touched[leaf] = True
    w0 = after[0::2]
    for d in range(1, D):
        nodes = np.unique(path[d - 1])
        touched[nodes] = True
        nodes = nodes[nodes != 0]
        tiles = (w0[nodes] & MASK).astype(np.int64)
        want = average8(aw1[tiles[:, None] + np.arange(8)[None, :]])
        assert np.array_equal(want, aw1[nodes]), "depth-%d node values are not averageChildren of their tiles" % d
    if canonical and n_after >= 8 and D >= 2 and ukeys.size:
        # Q6 canon: the root average is taken once, from node 0's own (pre-clobber) value: averageChildren of its
        # tile when node 0 lies on a touched path, else the word it held before the frame
        v = aw1[:8].copy()
        v[0] = bw1[0]
        if touched[0] and (w0[0] & FLAG):
            t0 = int(w0[0] & MASK)
            v[0] = average8(aw1[t0:t0 + 8][None, :])[0]
        assert aw1[0] == average8(v[None, :])[0], "node 0 (Q6) is not the canonical root average"
    untouched = ~touched
    untouched[0] = False
    assert np.array_equal(aw1[untouched], bw1[untouched]), "a value word off the touched paths changed"
    return ukeys.size
