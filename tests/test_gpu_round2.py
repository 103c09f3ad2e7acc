"""GPU: round-2 additions to the integrate path -- zero-copy colour gather for pinned host frames, reader -> writer
ordering around pipelined frames, pool validation on upload."""
import numpy as np
import pytest

from common import pkg, view_for_pose
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    return pkg()


@pytest.mark.parametrize("res", [(160, 120), (37, 23), (64, 31)])
@pytest.mark.parametrize("zero_copy", [True, False])
def test_pinned_host_frames_match_oracle(P, res, zero_copy):
    """osl_integrate_depth_host with PINNED planes: the colour plane is not copied, k_levels reads the winners in place
    (one or two aligned 32-bit loads; odd sizes exercise the end-of-buffer guard).  Same pool as the oracle, and as the
    staged path (`zero_copy=False`, what pageable planes get)."""
    import torch
    w, h = res
    D = 8
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D, zero_copy=zero_copy)
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for k in range(5):
        pose = P.synth.orbit_pose(12 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        hd, hc = torch.from_numpy(depth).pin_memory(), torch.from_numpy(rgb).pin_memory()
        keep.append((hd, hc))
        svo.integrate_depth_host(hd.numpy(), hc.numpy(), fx, fy, pose)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
    assert svo.size == ref.size
    assert np.array_equal(svo.pool(), ref.pool())


def test_raycast_between_pipelined_frames_sees_a_whole_map(P):
    """integrate(f); raycast(stream R); integrate(f+1): frame f+1's pool-writing stages run on the library's own
    streams and must wait for the raycast queued before them (readers were only ordered after writers).
    Every image must equal the one a strict replay renders after frame f."""
    import torch
    D, w, h = 10, 320, 240
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    n = 10
    frames = []
    for k in range(n):
        pose = P.synth.orbit_pose(15 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        frames.append((torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), pose))
    torch.cuda.synchronize()
    RW, RH = 1920, 1080  # a long reader: ~1 ms against ~30 us per frame
    R = torch.cuda.Stream()
    S = torch.cuda.Stream()
    piped = P.SVO(center, half, D).set_pipeline(True)
    outs = [torch.zeros((RH, RW, 4), dtype=torch.uint8, device="cuda") for _ in range(n)]
    for k, (d, c, pose) in enumerate(frames):
        piped.integrate_depth(d, c, fx, fy, pose, stream=S.cuda_stream)
        piped.raycast_device(outs[k], RW, RH, 45.0, view_for_pose(pose), stream=R.cuda_stream)
    torch.cuda.synchronize()
    strict = P.SVO(center, half, D)
    want = torch.zeros((RH, RW, 4), dtype=torch.uint8, device="cuda")
    for k, (d, c, pose) in enumerate(frames):
        strict.integrate_depth(d, c, fx, fy, pose)
        strict.raycast_device(want, RW, RH, 45.0, view_for_pose(pose))
        torch.cuda.synchronize()
        assert torch.equal(outs[k], want), "image after frame %d is torn (%d pixels differ)" % (
            k, int((outs[k] != want).any(dim=2).sum()))
    assert np.array_equal(piped.pool(), strict.pool())


def test_foreign_reader_on_a_joined_stream_is_ordered_before_later_frames(P):
    """osl_svo_join(stream) + foreign work on that stream (here: a device-to-device copy of the pool) + the next
    pipelined frame: the copy must see the pool after frame f, whole."""
    import torch
    D, w, h = 9, 320, 240
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D, reserve_nodes=1 << 21).set_pipeline(True)
    strict = P.SVO(center, half, D)
    F = torch.cuda.Stream()
    frames = []
    for k in range(6):
        pose = P.synth.orbit_pose(20 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        frames.append((torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), pose))
    torch.cuda.synchronize()
    d, c, pose = frames[0]
    svo.integrate_depth(d, c, fx, fy, pose)
    ptr, n0, _, _ = svo.view()  # synchronises; the pool does not move afterwards (reserve is large enough)

    class _Alias:  # the pool as a torch tensor (no copy): __cuda_array_interface__ over the raw device pointer
        __cuda_array_interface__ = {"shape": (1 << 22,), "typestr": "<i4", "data": (ptr, False), "version": 2}

    pool_t = torch.as_tensor(_Alias(), device="cuda")
    snaps = []
    for k in range(1, 6):
        d, c, pose = frames[k]
        svo.join(F.cuda_stream)
        snap = torch.empty(1 << 22, dtype=torch.int32, device="cuda")  # 2 words x 2^21 nodes: the whole reserve
        with torch.cuda.stream(F):
            snap.copy_(pool_t, non_blocking=True)  # foreign work: torch's own copy kernel on stream F
        snaps.append(snap)
        svo.integrate_depth(d, c, fx, fy, pose)
    torch.cuda.synchronize()
    d, c, pose = frames[0]
    strict.integrate_depth(d, c, fx, fy, pose)
    for k in range(1, 6):
        want = strict.pool()
        got = snaps[k - 1].cpu().numpy().view(np.uint32)
        assert np.array_equal(got[:want.size], want), "snapshot before frame %d is torn" % k
        assert not got[want.size:].any(), "snapshot before frame %d contains nodes of a later frame" % k
        d, c, pose = frames[k]
        strict.integrate_depth(d, c, fx, fy, pose)


def test_upload_rejects_corrupt_child_pointers(P):
    D = 6
    svo = P.SVO((0, 0, 0), 1.0, D)
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.9, 0.9, size=(3000, 3)).astype(np.float32)
    svo.integrate_points(pts, rng.integers(0, 256, size=(3000, 3)).astype(np.uint8))
    good = svo.pool()
    other = P.SVO((0, 0, 0), 1.0, D)
    other.load(good)
    assert np.array_equal(other.pool(), good)
    for poison in (good.size // 2 + 64, 12, 9):  # beyond the pool / inside the root tile / not 8-aligned
        bad = good.copy()
        victim = int(np.flatnonzero(bad[0::2] & 0x40000000)[3])
        bad[2 * victim] = 0x40000000 | poison
        with pytest.raises(P.OslError):
            other.load(bad)
        assert other.size == 0  # an empty, valid tree is left behind
    other.load(good)
    assert np.array_equal(other.raycast(64, 48), svo.raycast(64, 48))


@pytest.mark.parametrize("host_frames", [False, True])
@pytest.mark.parametrize("D", [8, 16])
def test_one_launch_per_frame_pipeline_matches_oracle(P, D, host_frames):
    """k_frame (one launch per frame: emit of frame f, sort of f-1, structure of f-2, values of f-3 as roles of one grid, chained
    by programmatic dependent launch) against the oracle AND against the same frames run as four kernels: pools
    bit-identical after every few frames, with a raycast, a size query and a scene cut in between (each of them drains
    or flushes the pipeline at a different depth)."""
    import torch
    w, h, n = 320, 240, 24
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    fused = P.SVO(center, half, D, reserve_nodes=1 << 22).set_pipeline(True)
    plain = P.SVO(center, half, D, reserve_nodes=1 << 22, fused=False).set_pipeline(True)
    ref = orc.OracleSVO(center, half, D)
    l0 = P.lib().osl_launch_count()
    keep = []
    for k in range(n):
        pose = P.synth.orbit_pose(4 * k if k < 16 else 400 + 4 * k)  # scene cut at frame 16: the splitters go stale
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
        for svo in (fused, plain):
            if host_frames:
                hd, hc = torch.from_numpy(depth).pin_memory(), torch.from_numpy(rgb).pin_memory()
                keep.append((hd, hc))
                svo.integrate_depth_host(hd.numpy(), hc.numpy(), fx, fy, pose)
            else:
                dd, cc = torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()
                keep.append((dd, cc))
                torch.cuda.synchronize()
                svo.integrate_depth(dd, cc, fx, fy, pose)
        if k in (5, 11, 12, 19):  # 12 right after 11: a flush with a single frame in the pipeline
            want = ref.pool()
            assert fused.size == ref.size and plain.size == ref.size
            assert np.array_equal(fused.pool(), want), "k_frame pipeline differs from the oracle after frame %d" % k
            assert np.array_equal(plain.pool(), want)
        if k == 8:
            img = fused.raycast(96, 72, 45.0, view_for_pose(pose))
            assert np.array_equal(img, ref.raycast(96, 72, 45.0, view_for_pose(pose)))
    assert np.array_equal(fused.pool(), ref.pool())
    assert np.array_equal(plain.pool(), ref.pool())
    c = fused.counters()
    assert c.frames == n and c.n_nodes == ref.size
    assert P.lib().osl_launch_count() > l0


def test_one_launch_per_frame_is_really_one_launch(P):
    """steady state of the pipelined path: one kernel launch per frame (+ three to drain)"""
    import torch
    D, w, h = 12, 320, 240
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D, reserve_nodes=1 << 22).set_pipeline(True)
    frames = []
    for k in range(40):
        pose = P.synth.orbit_pose(2 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        frames.append((torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), pose))
    torch.cuda.synchronize()
    for d, c, pose in frames[:10]:
        svo.integrate_depth(d, c, fx, fy, pose)
    svo.sync()
    l0 = P.lib().osl_launch_count()
    for d, c, pose in frames[10:]:
        svo.integrate_depth(d, c, fx, fy, pose)
    svo.sync()
    assert P.lib().osl_launch_count() - l0 == 30 + 3


@pytest.mark.parametrize("D", [10, 16, 20])
def test_walk_hints_are_only_guesses(P, D):
    """The structure stage walks keys that start at the root with the whole warp: one level per lane, nodes taken from a
    hint table and the links between the levels verified against the pool (walk_frontier).  A hint may be anything --
    here the table is overwritten with random words (valid node indices of OTHER nodes, indices beyond the pool,
    0xFFFFFFFF) before every frame, pipelined and strict: the pools stay equal to the oracle's, word for word."""
    import torch
    w, h, n = 320, 240, 12
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    piped = P.SVO(center, half, D, reserve_nodes=1 << 22).set_pipeline(True)
    strict = P.SVO(center, half, D, reserve_nodes=1 << 22)
    ref = orc.OracleSVO(center, half, D)
    keep = []
    for k in range(n):
        pose = P.synth.orbit_pose(3 * k if k < 8 else 300 + 3 * k)
        depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
        ref.integrate_depth(depth, rgb, fx, fy, pose)
        dd, cc = torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()
        keep.append((dd, cc))
        torch.cuda.synchronize()
        for svo in (piped, strict):
            if k % 2 == 1 or svo is strict:  # (the scramble synchronizes: every other frame the pipeline stays full)
                assert P.lib().osl_debug_scramble_hints(svo._h, 1000 * k + D) == 0
            svo.integrate_depth(dd, cc, fx, fy, pose)
    want = ref.pool()
    assert piped.size == ref.size and strict.size == ref.size
    assert np.array_equal(piped.pool(), want)
    assert np.array_equal(strict.pool(), want)
