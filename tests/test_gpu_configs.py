"""GPU: the MEASURED configurations themselves against the reference's own CUDA code (oracle/_ref, built from
/root/reference by oracle/Makefile) -- not scaled-down cousins ("pin the measured configurations"):

  cfg1   one 640x480 frame of scene S0 -> depth-8 SVO, 640x480 raycast: UNMODIFIED reference (D <= 10)
  bench  the map bench.py builds (640x480 orbit -> depth 16, pipelined): raycast at 640x480 and 1920x1080 against
         the reference's coneTraceSVO on the same (uploaded) pool
  cfg3   the 1000-frame 640x480 orbit into a depth-14 SVO, pipelined: node indices / child pointers against
         "ref + 64-bit patch" (the unmodified reference truncates keys beyond depth 10)

Where the reference races (Q7 duplicate keys, Q6 node 0) the comparison is on the set of legal outcomes, per leaf
(tests/common.check_frame_outcome); everything else is bit-exact."""
import concurrent.futures as cf
import multiprocessing as mp
import os
import time

import numpy as np
import pytest

from common import LOOK_PLUS_Z, check_frame_outcome, pkg, view_for_pose
from oracle import oracle as orc
from oracle import ref as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref/libosl_ref.so not built (needs /root/reference)")]

W, H = 640, 480


@pytest.fixture(scope="module")
def P():
    return pkg()


def _frame(args):
    k, seed = args
    from common import pkg as _pkg
    S = _pkg().synth
    pose = S.orbit_pose(k)
    d, c = S.make_frame(W, H, pose, seed=seed)
    return d, c, pose


def orbit_frames(n, seed0=0):
    """frames 0..n-1 of the cfg3 orbit (numpy ray-casting of scene S0; spread over the host cores)"""
    workers = max(1, min(16, (os.cpu_count() or 2) - 1))
    with cf.ProcessPoolExecutor(workers, mp_context=mp.get_context("fork")) as ex:  # the children only run numpy
        return list(ex.map(_frame, [(k, seed0 + k) for k in range(n)], chunksize=8))


def test_cfg1_frame_and_raycast_match_unmodified_reference(P):
    """SURVEY.md 8d cfg1, main.cpp:38-44 + cone_tracing_kernels.cu:157-198 of the UNMODIFIED reference."""
    D = 8
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(W, H)
    depth, rgb = P.synth.make_frame(W, H, None, seed=0)
    keys = orc.compute_keys(orc.vertex_map(depth, fx, fy), center, half, D)
    svo = P.SVO(center, half, D)
    ref = R.RefSVO(center, half, D)
    assert not ref.patched64
    empty = np.zeros(0, dtype=np.uint32)
    svo.integrate_depth(depth, rgb, fx, fy)
    ref.integrate_depth(depth, rgb, fx, fy)
    a, b = svo.pool(), ref.pool()
    assert svo.size == ref.size
    assert np.array_equal(a[0::2], b[0::2]), "node indices / child pointers differ from the reference"
    n_leaves = check_frame_outcome(empty, a, keys, rgb, D, canonical=True)
    assert check_frame_outcome(empty, b, keys, rgb, D, canonical=False) == n_leaves
    assert n_leaves == svo.counters().n_unique
    # the second observation of the same frame exercises Q3 (leaf splits) and the blend with a non-empty leaf
    svo.integrate_depth(depth, rgb, fx, fy)
    ref.integrate_depth(depth, rgb, fx, fy)
    a2, b2 = svo.pool(), ref.pool()
    assert svo.size == ref.size and svo.size > a.size // 2
    assert np.array_equal(a2[0::2], b2[0::2])
    check_frame_outcome(a, a2, keys, rgb, D, canonical=True)
    check_frame_outcome(b, b2, keys, rgb, D, canonical=False)
    # raycast 640x480 from the same pose: both kernels on BOTH pools (so the integrate races cannot hide anything)
    for pool in (a2, b2):
        svo.load(pool)
        ref.load(pool)
        for view in (np.eye(4, dtype=np.float32), LOOK_PLUS_Z):
            img = svo.raycast(W, H, 45.0, view)
            want, _ = ref.raycast(W, H, 45.0, view)
            assert np.array_equal(img, want), "%d pixels differ" % np.count_nonzero(np.any(img != want, axis=2))
    # after two observations alpha is 131: no sample terminates a ray early (Q8) and the picture is nearly black.  A map
    # observed 64 more times (alpha saturated) gives the renderer something to show; same comparison
    svo.load(a2)
    for _ in range(64):
        svo.integrate_depth(depth, rgb, fx, fy)
    ref.load(svo.pool())
    img = svo.raycast(W, H, 45.0, LOOK_PLUS_Z)
    want, _ = ref.raycast(W, H, 45.0, LOOK_PLUS_Z)
    assert np.array_equal(img, want), "%d pixels differ" % np.count_nonzero(np.any(img != want, axis=2))
    assert np.count_nonzero(img[..., :3].any(axis=2)) > W * H // 2  # the scene is in view


def test_bench_map_raycast_matches_reference_at_bench_sizes(P):
    """The raycast bench.py measures: the depth-16 map after 25 pipelined orbit frames, 640x480 and 1920x1080 from
    the last pose, against the reference's coneTraceSVO (ref + 64-bit patch library; cone_tracing_kernels.cu is
    unmodified in it) on the uploaded pool."""
    if not R.available(True):
        pytest.skip("libosl_ref64.so not built")
    D, n = 16, 25
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(W, H)
    frames = orbit_frames(n)
    svo = P.SVO(center, half, D, reserve_nodes=1 << 22).set_pipeline(True)
    keep = []
    import torch
    for d, c, pose in frames:
        dd, cc = torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()
        keep.append((dd, cc))
        torch.cuda.synchronize()  # pipelined mode: inputs complete at call time
        svo.integrate_depth(dd, cc, fx, fy, pose)
    pool = svo.pool()
    ref = R.RefSVO(center, half, D, patched64=True)
    ref.load(pool)
    view = view_for_pose(frames[-1][2])
    st = P.RaycastStats()
    img = svo.raycast(W, H, 45.0, view, stats=st)
    want, _ = ref.raycast(W, H, 45.0, view)
    assert np.array_equal(img, want), "%d pixels differ" % np.count_nonzero(np.any(img != want, axis=2))
    assert st.steps > 20 * st.rays  # a real march, not an empty map
    hd = svo.raycast(1920, 1080, 45.0, view)
    want_hd, _ = ref.raycast(1920, 1080, 45.0, view)
    assert np.array_equal(hd, want_hd), "%d pixels differ" % np.count_nonzero(np.any(hd != want_hd, axis=2))
    # the multi-GPU entry point renders the same pixels: interleaved bands of rank 1 of 4
    band = P.shard.band_height(1080, 4)
    rows = P.shard.row_bands(1080, 4, 1, band)
    out = torch.empty((sum(r for _, r in rows), 1920, 4), dtype=torch.uint8, device="cuda")
    svo.raycast_bands(out, 1920, 1080, band, 4, 1, 45.0, view)
    got = out.cpu().numpy()
    off = 0
    for row0, r in rows:
        assert np.array_equal(got[off:off + r], want_hd[row0:row0 + r])
        off += r


def test_cfg3_1000_frame_orbit_matches_patched_reference(P):
    """1000 frames 640x480 of S0 on the 0.36 deg/frame orbit into ONE depth-14 SVO, pipelined
    (host frames, the e2e path).  The reference (ref + 64-bit patch) is run first under a time budget -- its host-bound
    path ran 4 to 180 ms per frame on different boxes -- and ours is run over the same frames."""
    if not R.available(True):
        pytest.skip("libosl_ref64.so not built")
    D, n = 14, 1000
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(W, H)
    frames = orbit_frames(n)
    ref = R.RefSVO(center, half, D, patched64=True)
    t0, done = time.time(), 0
    for d, c, pose in frames:
        ref.integrate_depth(d, c, fx, fy, pose)
        done += 1
        if time.time() - t0 > 400.0:
            break
    svo = P.SVO(center, half, D, reserve_nodes=1 << 23)
    for d, c, pose in frames[:done]:
        svo.integrate_depth_host(d, c, fx, fy, pose)
    a, b = svo.pool(), ref.pool()
    assert svo.size == ref.size, (svo.size, ref.size, done)
    assert np.array_equal(a[0::2], b[0::2]), "node indices / child pointers differ after %d frames" % done
    # values: alpha is the canonical minimum (one +2 per observed leaf per frame, saturating), so ours <= reference
    aa, ab = (a[1::2] >> 24).astype(np.int32), (b[1::2] >> 24).astype(np.int32)
    assert np.all(aa <= ab)
    assert np.array_equal(aa > 127, ab > 127), "the sets of observed nodes differ"
    assert done == n, "the reference integrated only %d of %d frames inside its time budget (pools equal so far)" % (done, n)


@pytest.mark.parametrize("D", [10, 12])
def test_cfg2_voxel_grid_matches_reference(P, D):
    """SURVEY.md 8d cfg2, the mesh path main.cpp runs once per scene (Scene::voxelizeMeshes -> svoFromVoxelGrid,
    scene.cpp:64-85, svo.cu:584-640): bunny_tex.obj voxelised on the cfg2 cube, the whole grid (3.6 M voxels at depth 10,
    57.5 M at depth 12 -- the size bench.py's cfg2 object times) through OUR big-input kernels and through the
    reference's own svoFromVoxelGrid: depth 10 against the UNMODIFIED reference, depth 12 against ref + 64-bit patch
    (keys beyond depth 10 are truncated in the original, Q2).  Build + re-observation (Q3 leaf splits) + steady state,
    in Morton order and shuffled.  The voxeliser emits distinct cells, so nothing races: every word is compared (node 0's
    value apart, Q6)."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as graft
    if D > 10 and not R.available(True):
        pytest.skip("libosl_ref64.so not built")
    path = graft.asset("bunny_tex.obj")
    if path:
        V, T = P.synth.load_obj(path)
    else:  # (assets not staged: a mesh of the same size class)
        V, T = P.synth.icosphere(4, 1.35)
    colors = np.random.default_rng(2).uniform(0.0, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    lo, hi = V.min(axis=0), V.max(axis=0)
    center = tuple(float(x) for x in (np.float32(0.5) * (lo + hi)))
    half = float(hi[0])  # scene.cpp:78
    cen, col = P.meshToVoxelGrid(V, T, colors, center, half, D)
    n = cen.shape[0]
    assert n > (1 << 20)
    for order in ("morton", "shuffled"):
        if order == "shuffled":
            perm = torch.from_numpy(np.random.default_rng(4).permutation(n)).cuda()
            cen, col = cen[perm].contiguous(), col[perm].contiguous()
            del perm
        cen_h, col_h = cen.cpu().numpy(), col.cpu().numpy()
        svo = P.SVO(center, half, D, reserve_nodes=int(3.8 * n))
        ref = R.RefSVO(center, half, D, patched64=D > 10)
        for k in range(3):
            svo.integrate_voxels(cen, col)
            ref.integrate_voxels(cen_h, col_h)
            assert svo.size == ref.size, (order, k, svo.size, ref.size)
        a, b = svo.pool(), ref.pool()
        assert np.array_equal(a[0::2], b[0::2]), "%s: child pointers differ" % order
        a[1] = b[1] = 0
        assert np.array_equal(a, b), "%s: %d values differ" % (order, np.count_nonzero(a != b))
        svo.close()
        del ref, a, b
