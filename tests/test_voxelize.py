"""Sparse mesh voxeliser (osl_voxelize_mesh; the feeder of svoFromVoxelGrid for BASELINE configs 2 and 5).
The reference's voxeliser does not build (voxelpipe, CUDA 12), so the oracle here is the CPU statement of OUR contract
(oracle/osl_oracle.c: orc_voxelize_mesh, brute force over bounding boxes): parity with the reference is unpinned for
this function and bit-exact between GPU and oracle."""
import os

import numpy as np
import pytest

from common import LOOK_PLUS_Z, check_pool_invariants, pkg
from oracle import oracle as orc


def _decode(keys, D):
    ix = np.zeros(keys.size, dtype=np.int64)
    iy, iz = ix.copy(), ix.copy()
    for l in range(D - 1, -1, -1):
        d = (keys >> (3 * l)) & 7
        ix = (ix << 1) | (d & 1)
        iy = (iy << 1) | ((d >> 1) & 1)
        iz = (iz << 1) | ((d >> 2) & 1)
    return ix, iy, iz


def test_oracle_single_triangle_known_cells():
    # an axis-aligned right triangle in the plane z = 0.3 of the cube [-1,1]^3 at depth 3 (cells of 0.25)
    V = np.array([[-0.9, -0.9, 0.3], [0.4, -0.9, 0.3], [-0.9, 0.4, 0.3]], dtype=np.float32)
    T = np.array([[0, 1, 2]], dtype=np.int32)
    keys, tris, cen = orc.voxelize_mesh(V, T, (0, 0, 0), 1.0, 3)
    ix, iy, iz = _decode(keys, 3)
    assert np.all(iz == 5) and np.all(tris == 0)          # z = 0.3 lies in cell 5 of [-1, 1] / 0.25
    assert np.all(np.diff(keys) > 0)
    got = set(zip(ix.tolist(), iy.tolist()))
    want = set()
    for x in range(8):
        for y in range(8):
            # cell [x0,x0+.25]x[y0,y0+.25] overlaps the triangle x >= -0.9, y >= -0.9, x + y <= -0.5
            x0, y0 = -1 + 0.25 * x, -1 + 0.25 * y
            if x0 + 0.25 >= -0.9 and y0 + 0.25 >= -0.9 and x0 + y0 <= -0.5 and x0 <= 0.4 and y0 <= 0.4:
                want.add((x, y))
    assert got == want
    assert np.allclose(cen[:, 2], -1 + 5.5 * 0.25) and np.all(cen[:, 3] == 1.0)


def test_oracle_sphere_shell_is_closed_and_thin():
    P = pkg()
    V, T = P.synth.icosphere(2, 0.7)
    D = 5
    keys, tris, cen = orc.voxelize_mesh(V, T, (0, 0, 0), 1.0, D)
    r = np.linalg.norm(cen[:, :3], axis=1)
    cs = 2.0 / 2 ** D
    assert np.all(np.abs(r - 0.7) < 0.05 + cs)   # every voxel touches the (faceted) sphere
    assert keys.size > 4 * np.pi * 0.7 ** 2 / cs ** 2 * 0.9  # and the shell has no holes to speak of
    assert tris.min() >= 0 and tris.max() < T.shape[0]


@pytest.mark.skipif(not os.path.exists("/root/reference/objs/bunny_tex.obj"), reason="reference assets not mounted")
def test_oracle_voxelises_the_reference_bunny():
    P = pkg()
    V, T = P.synth.load_obj("/root/reference/objs/bunny_tex.obj")
    assert T.shape[0] == 4968
    lo, hi = V.min(axis=0), V.max(axis=0)
    center = (lo + hi) / 2
    half = float((hi - lo).max() / 2 * 1.01)
    keys, tris, cen = orc.voxelize_mesh(V, T, tuple(center), half, 6)
    assert 2000 < keys.size < 64 ** 3 // 4 and np.all(np.diff(keys) > 0)


# ------------------------------------------------------------------------------------------------ GPU
def _random_soup(rng, n):
    V = rng.uniform(-1.2, 1.2, size=(3 * n, 3)).astype(np.float32)
    T = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    V[3] = V[4]                       # degenerate (edge)
    V[6] = V[7] = V[8]                # degenerate (point)
    V[9:12] += 5.0                    # entirely outside the cube
    V[12:15, 2] = 0.25                # exactly on a cell boundary plane
    small = rng.integers(5, n, size=n // 2)
    for t in small:                   # many small triangles
        c = rng.uniform(-0.9, 0.9, size=3)
        V[3 * t:3 * t + 3] = (c + rng.normal(scale=0.03, size=(3, 3))).astype(np.float32)
    return V, T


@pytest.mark.gpu
@pytest.mark.parametrize("mesh,D", [("ico3", 6), ("ico3", 9), ("ico1", 4), ("soup", 5), ("soup", 7), ("big", 8)])
def test_voxelize_matches_oracle_bit_exact(mesh, D):
    P = pkg()
    rng = np.random.default_rng(D)
    if mesh == "ico3":
        V, T = P.synth.icosphere(3, 0.8, (0.05, -0.02, 0.1))
    elif mesh == "ico1":
        V, T = P.synth.icosphere(1, 0.9)
    elif mesh == "soup":
        V, T = _random_soup(rng, 300)
    else:  # two huge triangles (a floor) + a small one: chunked column traversal
        V = np.array([[-3, -0.31, -3], [3, -0.31, -3], [3, -0.31, 3], [-3, -0.31, 3], [0, 0, 0], [0.1, 0.2, 0], [0, 0.2, 0.3]],
                     dtype=np.float32)
        T = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], dtype=np.int32)
    colors = rng.uniform(0, 1, size=(T.shape[0], 4)).astype(np.float32)
    center, half = (0.0, 0.0, 0.0), 1.0
    c, k, keys, tris = P.meshToVoxelGrid(V, T, colors, center, half, D, want_keys=True)
    wk, wt, wc = orc.voxelize_mesh(V, T, center, half, D)
    assert keys.shape[0] == wk.size
    assert np.array_equal(keys.cpu().numpy(), wk)
    assert np.array_equal(tris.cpu().numpy(), wt)
    assert np.array_equal(c.cpu().numpy().view(np.uint32), wc.view(np.uint32))
    assert np.array_equal(k.cpu().numpy(), colors[wt])


@pytest.mark.gpu
def test_mesh_to_svo_to_image_cfg2_shape():
    """BASELINE configs[1] shape on a procedural mesh: voxelise -> svoFromVoxelGrid -> raycast, against the oracle"""
    P = pkg()
    D = 8
    V, T = P.synth.icosphere(3, 0.8)
    rng = np.random.default_rng(1)
    colors = rng.uniform(0.2, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    center, half = (0.0, 0.0, 0.0), 1.0
    cen, col = P.meshToVoxelGrid(V, T, colors, center, half, D)
    svo = P.SVO(center, half, D)
    ref = orc.OracleSVO(center, half, D)
    cen_h, col_h = cen.cpu().numpy(), col.cpu().numpy()
    for _ in range(66):  # saturate alpha so that the shell is opaque to the raycaster
        svo.integrate_voxels(cen, col)
    for _ in range(66):
        ref.integrate_voxels(cen_h, col_h)
    assert svo.size == ref.size
    pool = svo.pool()
    assert np.array_equal(pool, ref.pool())
    check_pool_invariants(pool)
    # Q11 does not scramble colours here: the grid is already in key order, so every leaf carries its own colour
    view = np.eye(4, dtype=np.float32)
    view[2, 3] = -2.5  # camera 2.5 half-edges back on +z, looking down -z (the reference renderer's convention)
    img = svo.raycast(192, 108, 45.0, view)
    want = orc.raycast(pool, center, half, 192, 108, 45.0, view)
    assert np.array_equal(img, want)
    assert np.count_nonzero(img[..., :3].sum(axis=2)) > 500  # the sphere is in the picture


# ---- the reference's own rule (voxelpipe THIN_RASTER on the dense grid over the mesh bounding box) -----------------

def _thin_centers(cells, bbox0, bbox1, log_n):
    """getCenterFromIndex (voxelization.cu:58-78) in float32: tiles of 8 voxels, M = N / 8 tiles per axis"""
    f = np.float32
    b0, b1 = np.asarray(bbox0, dtype=f), np.asarray(bbox1, dtype=f)
    M = f((1 << log_n) // 8)
    td = (b1 - b0) / M
    pd = td / f(8)
    t, p = (cells // 8).astype(f), (cells % 8).astype(f)
    return (b0[None, :] + t * td[None, :] + p * pd[None, :] + (pd / f(2.0))[None, :]).astype(f)


def test_thin_oracle_analytic_and_bunny():
    """CPU: the restated rule on hand-checkable input, and on the reference's bunny at the reference's 256^3."""
    # an axis-aligned triangle in the plane z = 0.3 of the unit cube at 16^3: exactly one voxel layer (w = int(0.3*16) =
    # 4), and -- 2-D conservative coverage -- every pixel the triangle's footprint TOUCHES
    V = np.array([[0.1, 0.1, 0.3], [0.9, 0.15, 0.3], [0.2, 0.85, 0.3]], dtype=np.float32)
    T = np.array([[0, 1, 2]], dtype=np.int32)
    cells, tris = orc.voxelize_thin(V, T, (0, 0, 0), (1, 1, 1), 4)
    assert np.all(cells[:, 2] == 4) and np.all(tris == 0)
    got = {(int(x), int(y)) for x, y, _ in cells}
    # brute force: pixel [x, x+1] x [y, y+1] (in cell units) overlaps the triangle <=> no separating edge line
    P2 = V[:, :2].astype(np.float64) * 16
    want = set()
    for x in range(16):
        for y in range(16):
            inside = True
            for i in range(3):
                a, b = P2[i], P2[(i + 1) % 3]
                nx, ny = -(b[1] - a[1]), (b[0] - a[0])  # inward normal for this winding (ccw in xy)
                cx, cy = (x + 1 if nx > 0 else x), (y + 1 if ny > 0 else y)  # the pixel corner furthest inside
                if nx * (cx - a[0]) + ny * (cy - a[1]) < 0:
                    inside = False
            if inside:
                want.add((x, y))
    assert got == want and len(got) > 40
    if not os.path.exists(BUNNY):
        pytest.skip("reference assets are not mounted")
    P = pkg()
    Vb, Tb = P.synth.load_obj(BUNNY)
    lo, hi = Vb.min(axis=0), Vb.max(axis=0)
    cells, tris = orc.voxelize_thin(Vb, Tb, lo, hi, 8)
    assert 100_000 < cells.shape[0] < 400_000  # a 2-manifold of area ~ 2.1 * 256^2 cells
    # every voxel lies within a cell of its triangle's plane and inside the triangle's (1-cell dilated) bounding box
    cen = _thin_centers(cells, lo, hi, 8).astype(np.float64)
    a, b, c = (Vb[Tb[tris][:, i]].astype(np.float64) for i in range(3))
    n = np.cross(b - a, c - a)
    n /= np.linalg.norm(n, axis=1)[:, None]
    cell = (hi - lo).astype(np.float64) / 256.0
    dist = np.abs(np.einsum("ij,ij->i", cen - a, n))
    assert dist.max() <= 0.87 * np.linalg.norm(cell)  # at most ~half a cell diagonal (+ the conservative 2-D margin)
    # (a column whose pixel only touches the footprint samples the plane up to a pixel outside the triangle)
    tlo, thi = np.minimum(np.minimum(a, b), c) - 3.5 * cell, np.maximum(np.maximum(a, b), c) + 3.5 * cell
    assert np.all(cen >= tlo) and np.all(cen <= thi)
    # watertight enough for a 6-separating voxelisation: every column of the dominant-axis projection that a triangle's
    # interior covers holds a voxel -- checked through the voxel count per triangle area (>= projected area in cells)
    assert np.unique(tris).size > 0.95 * Tb.shape[0]


BUNNY = "/root/reference/objs/bunny_tex.obj"
if not os.path.exists(BUNNY):
    _staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_assets", "bunny_tex.obj")
    if os.path.exists(_staged):
        BUNNY = _staged


@pytest.mark.gpu
@pytest.mark.parametrize("mesh,log_n", [("ico3", 6), ("soup", 5), ("soup", 7), ("big", 8), ("bunny", 8)])
def test_thin_voxeliser_matches_restated_reference_rule(mesh, log_n):
    """osl_voxelize_thin (CUDA) against oracle/osl_oracle_thin.c, the CPU restatement of the voxelpipe rule the
    reference calls (voxelization.cu:281-285): the same cells, the same lowest triangle per cell, the reference's
    voxel centres (getCenterFromIndex) bit for bit -- on the reference's bunny at the reference's 256^3 among others."""
    P = pkg()
    rng = np.random.default_rng(log_n)
    if mesh == "ico3":
        V, T = P.synth.icosphere(3, 0.8, (0.05, -0.02, 0.1))
    elif mesh == "soup":
        V, T = _random_soup(rng, 300)
    elif mesh == "big":
        V = np.array([[-3, -0.31, -3], [3, -0.31, -3], [3, -0.31, 3], [-3, -0.31, 3], [0, 0, 0], [0.1, 0.2, 0], [0, 0.2, 0.3]],
                     dtype=np.float32)
        T = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], dtype=np.int32)
    else:
        if not os.path.exists(BUNNY):
            pytest.skip("bunny_tex.obj not staged (baseline/_assets, __graft_entry__.build)")
        V, T = P.synth.load_obj(BUNNY)
    lo, hi = V.min(axis=0), V.max(axis=0)
    if mesh == "big":
        lo, hi = np.array([-1, -1, -1], np.float32), np.array([1, 1.5, 1], np.float32)  # triangles stick out of the box
    colors = rng.uniform(0, 1, size=(T.shape[0], 4)).astype(np.float32)
    cen, col, cells, tris = P.meshToVoxelGridThin(V, T, colors, lo, hi, log_n)  # grid order: z, y, x
    want_cells, want_tris = orc.voxelize_thin(V, T, lo, hi, log_n)
    assert cells.shape[0] == want_cells.shape[0] > 0
    assert np.array_equal(cells.cpu().numpy(), want_cells)
    assert np.array_equal(tris.cpu().numpy(), want_tris)
    assert np.array_equal(cen.cpu().numpy()[:, :3].view(np.uint32), _thin_centers(want_cells, lo, hi, log_n).view(np.uint32))
    assert np.array_equal(col.cpu().numpy(), colors[want_tris])
    # ordered for an octree cube (Scene::voxelizeMeshes' cube): same voxels, ascending Morton keys of their centres
    center = tuple(float(x) for x in (np.float32(0.5) * (lo + hi)))
    half = float(hi[0])
    cen2, col2, cells2, tris2 = P.meshToVoxelGridThin(V, T, colors, lo, hi, log_n, cube=(center, half, log_n))
    c2 = cells2.cpu().numpy()
    assert {tuple(r) for r in c2} == {tuple(r) for r in want_cells}
    keys = orc.compute_keys(cen2.cpu().numpy()[:, :3], center, half, log_n)
    assert np.all(np.diff(keys) >= 0)
