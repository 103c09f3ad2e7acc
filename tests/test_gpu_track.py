"""GPU parity tests of the camera-tracking path (osl_track.cu through the C ABI) against the CPU oracle and, where
oracle/_ref was built, against the reference's OWN CUDA kernels and its own RGBDCamera run on this GPU.

Tolerances (north_star: bit-exact integer work, 1e-4 on accumulated float values):
  * subsampleDepth, normal map, normal transform, intensity, subsample: bit-exact against both;
  * bilateral filter: bit-exact against the reference's kernel (same MUFU.EX2 sequence); against the CPU oracle
    (exp2f instead of MUFU.EX2) at most 1 mm on at most 0.5 % of the pixels;
  * ICP normal equations: float sums in a different (deterministic) order -> relative 1e-4 of the matrix norm;
  * tracked pose: 1e-4 absolute on the orientation entries."""
import numpy as np
import pytest

from common import float_bits_equal, pkg
from oracle import oracle as orc
from oracle import ref

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def P():
    return pkg()


def _depth(P, w, h, k=0, **kw):
    pose = P.synth.orbit_pose(5 * k)
    return P.synth.make_frame(w, h, pose, seed=k, **kw)[0]


def _poses(n):
    out = []
    for k in range(n):
        M = np.eye(4)
        a = np.radians(0.3 * k)
        M[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        M[:3, 3] = [0.01 * k, 0.0, 0.005 * k]
        out.append(M.astype(np.float32))
    return out


@pytest.mark.parametrize("w,h", [(640, 480), (333, 77), (32, 8), (9, 1)])
def test_bilateral_matches_oracle(P, w, h):
    d = _depth(P, w, h, 1)
    d[::7, ::5] = 0
    d[3 % h, 5 % w] = 65535
    got = P.sensor.bilateralFilter(d).cpu().numpy()
    want = orc.bilateral(d)
    diff = np.abs(got.astype(np.int64) - want.astype(np.int64))
    assert diff.max() <= 1, diff.max()
    assert np.mean(diff != 0) <= 0.005


@needs_ref
@pytest.mark.parametrize("w,h", [(640, 480), (333, 77)])
def test_bilateral_bit_exact_against_the_reference_kernel(P, w, h):
    d = _depth(P, w, h, 2)
    d[::9, ::4] = 0
    assert np.array_equal(P.sensor.bilateralFilter(d).cpu().numpy(), ref.bilateral(d))


@pytest.mark.parametrize("w,h", [(640, 480), (320, 240), (10, 6), (2, 2)])
def test_subsample_depth_bit_exact(P, w, h):
    d = _depth(P, w, h, 3)
    d[::5, ::3] = 0
    got = P.sensor.subsampleDepth(d).cpu().numpy()
    assert np.array_equal(got, orc.subsample_depth(d))
    if ref.available():
        assert np.array_equal(got, ref.subsample_depth(d))


@pytest.mark.parametrize("w,h", [(640, 480), (77, 33), (2, 2), (1, 1)])
def test_normal_map_bit_exact(P, w, h):
    d = _depth(P, w, h, 4)
    fx, fy = P.synth.focal(w, h)
    vtx = P.generateVertexMap(d, fx, fy)
    got = P.sensor.generateNormalMap(vtx, w, h).cpu().numpy()
    want = orc.normal_map(vtx.cpu().numpy(), w, h)
    assert float_bits_equal(got, want)
    if ref.available():
        assert float_bits_equal(got, ref.normal_map(vtx.cpu().numpy(), w, h))


def test_transform_normals_intensity_subsample_bit_exact(P):
    import torch
    rng = np.random.default_rng(11)
    n = rng.normal(size=(5000, 3)).astype(np.float32)
    n[::13] = np.inf
    M = _poses(3)[2]
    got = P.sensor.transformNormalMap(torch.from_numpy(n).cuda(), M).cpu().numpy()
    assert float_bits_equal(got, orc.transform_normals(n, M))
    rgb = rng.integers(0, 256, size=(4000, 3), dtype=np.uint8)
    gi = P.sensor.colorToIntensity(rgb).cpu().numpy()
    assert float_bits_equal(gi, orc.color_to_intensity(rgb))
    img = rng.normal(size=(48, 64)).astype(np.float32)
    gs = P.sensor.subsample(img).cpu().numpy()
    assert np.array_equal(gs, orc.subsample_f32(img))
    if ref.available():
        assert float_bits_equal(got, ref.transform_normals(n, M))
        assert float_bits_equal(gi, ref.color_to_intensity(rgb))
        assert np.array_equal(gs, ref.subsample_f32(img))


def _maps(P, w, h, k, pose):
    import torch
    d = P.synth.make_frame(w, h, pose, seed=k, invalid_frac=0.02, noise_mm=1)[0]
    fx, fy = P.synth.focal(w, h)
    f = P.sensor.bilateralFilter(d)
    v = P.generateVertexMap(f, fx, fy)
    n = P.sensor.generateNormalMap(v, w, h)
    torch.cuda.synchronize()
    return v, n


@pytest.mark.parametrize("w,h", [(640, 480), (160, 120), (41, 23)])
@pytest.mark.parametrize("exact", [False, True])
def test_icp_cost_matches_oracle(P, w, h, exact):
    poses = _poses(2)
    v1, n1 = _maps(P, w, h, 0, poses[0])
    v2, n2 = _maps(P, w, h, 1, poses[1])
    A, b, pairs = P.sensor.computeICPCost2(v1, n1, v2, n2, exact_jacobian=exact)
    rA, rb, rpairs = orc.icp_cost(v1.cpu().numpy(), n1.cpu().numpy(), v2.cpu().numpy(), n2.cpu().numpy(), exact)
    assert pairs == rpairs and pairs > 0.3 * w * h
    assert np.allclose(A, A.T)
    assert np.abs(A - rA).max() <= 1e-4 * np.abs(rA).max()
    assert np.abs(b - rb).max() <= 1e-4 * max(np.abs(rb).max(), 1e-3 * np.abs(rA).max())


@needs_ref
def test_icp_cost_matches_the_reference_kernel(P):
    w, h = 640, 480
    poses = _poses(2)
    v1, n1 = _maps(P, w, h, 0, poses[0])
    v2, n2 = _maps(P, w, h, 1, poses[1])
    A, b, _ = P.sensor.computeICPCost2(v1, n1, v2, n2)
    rA, rb = ref.icp_cost2(v1.cpu().numpy(), n1.cpu().numpy(), v2.cpu().numpy(), n2.cpu().numpy(), w, h)
    assert np.abs(A - rA).max() <= 1e-4 * np.abs(rA).max()
    assert np.abs(b - rb).max() <= 1e-4 * max(np.abs(rb).max(), 1e-3 * np.abs(rA).max())


def test_icp_cost_is_deterministic(P):
    w, h = 320, 240
    poses = _poses(2)
    v1, n1 = _maps(P, w, h, 0, poses[0])
    v2, n2 = _maps(P, w, h, 1, poses[1])
    first = P.sensor.computeICPCost2(v1, n1, v2, n2)
    for _ in range(3):
        again = P.sensor.computeICPCost2(v1, n1, v2, n2)
        assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])


@pytest.mark.parametrize("w,h,exact,host", [(320, 240, False, True), (320, 240, True, False), (640, 480, False, False),
                                            (640, 480, True, True)])
def test_tracker_matches_oracle(P, w, h, exact, host):
    import torch
    fx, fy = P.synth.focal(w, h)
    cam = P.RGBDCamera(w, h, (fx, fy), exact_jacobian=exact)
    want = orc.OracleTracker(w, h, fx, fy, exact_jacobian=exact)
    for k, pose in enumerate(_poses(4)):
        d = P.synth.make_frame(w, h, pose, seed=k, invalid_frac=0.01, noise_mm=1)[0]
        cam.update(d if host else torch.from_numpy(d).cuda())
        want.update(d)
        assert cam.lost == want.lost
        assert abs(cam.pairs - want.pairs) <= 0.002 * w * h  # 1 mm bilateral differences move a few pairs
        assert np.abs(cam.pose() - want.pose()).max() <= 1e-4, (k, cam.pose(), want.pose())
    if exact:
        truth = np.linalg.inv(_poses(4)[0]) @ _poses(4)[3]
        assert np.abs(cam.pose() - truth).max() <= 8e-3   # against the TRUE motion (1.6e-2 rad, 3.4e-2 m)
    else:
        assert np.array_equal(cam.position(), np.zeros(3, np.float32))  # quirk Q18
    # the pyramid of the last frame: vertex maps are exact up to the bilateral's rare 1 mm steps
    v, n = cam.level(2)
    assert v.shape == ((w // 4) * (h // 4), 3) and np.isfinite(v).mean() > 0.9


@needs_ref
def test_tracker_matches_the_reference_rgbd_camera(P):
    """the reference's own RGBDCamera::update (rgbd_camera.cpp) on this GPU against osl_tracker_update"""
    w, h = 640, 480
    fx, fy = P.synth.focal(w, h)
    cam = P.RGBDCamera(w, h, (fx, fy))
    rcam = ref.RefTracker(w, h, fx, fy)
    for k, pose in enumerate(_poses(4)):
        d = P.synth.make_frame(w, h, pose, seed=k, invalid_frac=0.01, noise_mm=1)[0]
        cam.update(d)
        rcam.update(d)
        assert np.abs(cam.orientation() - rcam.orientation()).max() <= 1e-4, (k, cam.orientation(), rcam.orientation())
        assert np.abs(cam.position() - rcam.position()).max() <= 1e-4
    assert not np.array_equal(cam.orientation(), np.eye(3, dtype=np.float32))


def test_tracker_blank_frames_and_reset(P):
    cam = P.RGBDCamera(64, 48, (50.0, 50.0))
    z = np.zeros((48, 64), np.uint16)
    cam.update(z)
    assert not cam.lost
    cam.update(z)
    assert cam.lost and cam.pairs == 0
    assert np.array_equal(cam.orientation(), np.eye(3, dtype=np.float32))
    cam.reset()
    d = _depth(P, 64, 48, 1)
    cam.update(d)
    cam.update(d)
    assert not cam.lost and cam.pairs > 1000
    assert np.abs(cam.pose() - np.eye(4)).max() < 1e-5   # the same frame twice: no motion


def test_tracker_feeds_integration(P):
    """the SLAM loop of main.cpp:33-44 with the tracking line uncommented: poses from the tracker drive the map"""
    w, h, D = 320, 240, 10
    fx, fy = P.synth.focal(w, h)
    center, half = P.synth.tree_params(D)
    cam = P.RGBDCamera(w, h, (fx, fy), exact_jacobian=True)
    svo = P.SVO(center, half, D)
    ref_svo = orc.OracleSVO(center, half, D)
    track = orc.OracleTracker(w, h, fx, fy, exact_jacobian=True)
    for k, pose in enumerate(_poses(4)):
        d, c = P.synth.make_frame(w, h, pose, seed=k, invalid_frac=0.01, noise_mm=1)
        cam.update(d)
        est = cam.pose()
        svo.integrate_depth(d, c, fx, fy, est)
        ref_svo.integrate_depth(d, c, fx, fy, est)     # same estimated pose: the map stays bit-exact
        track.update(d)
        assert np.abs(est - track.pose()).max() <= 1e-4
    assert np.array_equal(svo.pool(), ref_svo.pool())


@pytest.mark.parametrize("piped", [False, True])
@pytest.mark.parametrize("exact", [False, True])
def test_tracked_integration_without_host_round_trip(P, piped, exact):
    """osl_tracker_update + osl_integrate_depth_posed on one stream: the pose never visits the host before it is used;
    reading it back AFTERWARDS and replaying the oracle with it gives the same pool bit for bit"""
    import torch
    w, h, D = 320, 240, 10
    fx, fy = P.synth.focal(w, h)
    center, half = P.synth.tree_params(D)
    cam = P.RGBDCamera(w, h, (fx, fy), exact_jacobian=exact)
    svo = P.SVO(center, half, D).set_pipeline(piped)
    ref_svo = orc.OracleSVO(center, half, D)
    frames = [P.synth.make_frame(w, h, pose, seed=k, invalid_frac=0.01, noise_mm=1) for k, pose in enumerate(_poses(5))]
    dev = [(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()) for d, c in frames]
    torch.cuda.synchronize()
    used = []
    for d, c in dev:
        svo.integrate_depth_tracked(d, c, fx, fy, cam)
        used.append(cam.pose())          # waits for the tracker only; read after the frame was queued
    for (d, c), pose in zip(frames, used):
        ref_svo.integrate_depth(d, c, fx, fy, pose)
    assert svo.size == ref_svo.size
    assert np.array_equal(svo.pool(), ref_svo.pool())
    if exact:
        assert np.abs(used[-1][:3, 3] - [0.04, 0.0, 0.02]).max() < 8e-3
