"""Camera-tracking oracle (oracle/osl_oracle_track.c; reference image_kernels.cu:104-321, localization_kernels.cu,
rgbd_camera.cpp:53-224).  The reference ships no tests or vectors for this path: the oracle is pinned by hand-derived
known answers here and, on the GPU box, by the reference's own kernels and RGBDCamera (tests/test_gpu_track.py)."""
import numpy as np

from common import pkg
from oracle import oracle as orc


def test_bilateral_known_answers():
    flat = np.full((12, 16), 1000, dtype=np.uint16)
    out = orc.bilateral(flat)
    assert np.array_equal(out[:-1], flat[:-1])      # a constant image is a fixed point ...
    # ... except the last row: its window [y-3, min(y+4, h-1)) never contains the row itself, still all 1000
    assert np.array_equal(out[-1], flat[-1])
    one = np.zeros((1, 9), dtype=np.uint16)         # h = 1: the window is empty, 0/0 = NaN -> 0
    one[:] = 500
    assert np.array_equal(orc.bilateral(one), np.zeros_like(one))
    step = np.full((12, 16), 1000, dtype=np.uint16)
    step[:, 8:] = 3000                              # a 2 m step: exp(-(2000^2) * 3.1e-4) == 0, edges are preserved
    assert np.array_equal(orc.bilateral(step), step)
    noisy = flat.copy()
    noisy[5, 5] = 1040
    out = orc.bilateral(noisy)
    assert 1000 < out[5, 5] < 1040 and out[5, 6] >= 1000 and out[0, 15] == 1000


def test_subsample_depth_known_answers():
    a = np.full((8, 12), 800, dtype=np.uint16)
    a[::2, ::2] = 810                               # the sampled pixels; all within sigma*3 = 120 of each other
    out = orc.subsample_depth(a)
    assert out.shape == (4, 6)
    # interior output (1,1): window rows 0-4, cols 0-4 (5x5): 9 samples of 810, 16 of 800 -> trunc(803.6)
    assert out[1, 1] == (9 * 810 + 16 * 800) // 25
    # last column x = 5: cols [8, min(13, 11)) = 8, 9, 10 -> 2 of 3 columns sampled
    assert out[1, 5] == int((3 * 2 * 810 + (15 - 6) * 800) / 15)
    far = a.copy()
    far[3, 3] = 5000                                # outside sigma of the centre: excluded from the mean
    assert orc.subsample_depth(far)[1, 1] == (9 * 810 + 15 * 800) // 24
    hole = a.copy()
    hole[2, 2] = 0                                  # an invalid centre averages only other near-zero values
    assert orc.subsample_depth(hole)[1, 1] == 0


def test_normal_map_of_a_plane():
    w, h = 6, 5
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    # image y grows downwards while camera y grows upwards (generateVertexMap): a fronto-parallel plane z = 2
    vtx = np.stack([xs * 0.01, -ys * 0.01, np.full_like(xs, 2.0)], axis=2).reshape(-1, 3)
    n = orc.normal_map(vtx, w, h).reshape(h, w, 3)
    assert np.all(np.isinf(n[:, -1])) and np.all(np.isinf(n[-1, :]))
    assert np.allclose(n[:-1, :-1], [0, 0, 1], atol=1e-6)    # -(right x down): away from the camera
    tilted = vtx.copy()
    tilted[:, 2] += tilted[:, 0]                    # z = 2 + x: normal ~ (-1, 0, 1) / sqrt(2)
    n = orc.normal_map(tilted, w, h).reshape(h, w, 3)
    assert np.allclose(n[0, 0], np.array([-1, 0, 1]) / np.sqrt(2), atol=1e-5)
    bad = vtx.copy()
    bad[7] = np.inf                                 # an invalid vertex poisons the normals that touch it
    n = orc.normal_map(bad, w, h).reshape(h, w, 3)
    assert not np.all(np.isfinite(n[1, 1])) and not np.all(np.isfinite(n[1, 0])) and not np.all(np.isfinite(n[0, 1]))
    assert np.all(np.isfinite(n[2, 2]))


def test_intensity_ignores_green_and_subsample_picks_even_pixels():
    rgb = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255]], dtype=np.uint8)
    got = orc.color_to_intensity(rgb)
    assert np.allclose(got, [0.299, 0.0, 0.587 + 0.114], atol=1e-6)  # quirk: r, b, b
    img = np.arange(48, dtype=np.float32).reshape(6, 8)
    assert np.array_equal(orc.subsample_f32(img), img[::2, ::2])


def test_cholesky_and_pose_increment():
    rng = np.random.default_rng(5)
    M = rng.normal(size=(6, 6))
    A = (M @ M.T + 6 * np.eye(6)).astype(np.float32)
    x = rng.normal(size=6).astype(np.float32)
    b = A @ x
    assert np.allclose(orc.solve_cholesky(A, b), x, atol=1e-4)
    assert np.all(np.isnan(orc.solve_cholesky(np.zeros((6, 6), np.float32), np.zeros(6, np.float32))))  # "lost"
    T = orc.pose_increment([0, 0, 0, 0.1, 0.2, 0.3])
    assert np.allclose(T, [[1, 0, 0, 0.1], [0, 1, 0, 0.2], [0, 0, 1, 0.3], [0, 0, 0, 1]])
    a = 0.01
    T = orc.pose_increment([0, 0, a, 0, 0, 0])      # the reference rotates by MINUS the solved angle
    assert np.allclose(T[:2, :2], [[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]], atol=1e-6)
    T = orc.pose_increment([0, 0, a, 1, 0, 0], exact_jacobian=True)
    assert np.allclose(T[:2, :2], [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]], atol=1e-6)
    assert np.allclose(T[:3, 3], [1, 0, 0])         # translate after rotating
    T = orc.pose_increment([0, 0, a, 1, 0, 0])
    assert np.allclose(T[:3, 3], [np.cos(a), -np.sin(a), 0], atol=1e-6)  # the reference rotates the translation


def test_icp_cost_single_pair_by_hand():
    v1 = np.array([[0.1, 0.2, 1.0]], np.float32)
    n1 = np.array([[0.0, 0.0, -1.0]], np.float32)
    v2 = np.array([[0.1, 0.2, 1.02]], np.float32)
    A, b, pairs = orc.icp_cost(v1, n1, v2, n1)
    assert pairs == 1
    at = np.array([-0.1 * 0 - 0.2 * -1, -1.02 * 0 + 0.1 * -1, 0.2 * 0 + 1.02 * 0, 0, 0, -1], np.float32)  # Q17 rows
    assert np.allclose(A, np.outer(at, at), atol=1e-7)
    assert np.allclose(b, 0.02 * at, atol=1e-7)     # n . (v1 - v2) = 0.02
    A, b, _ = orc.icp_cost(v1, n1, v2, n1, exact_jacobian=True)
    at = np.concatenate([np.cross(v2[0], n1[0]), n1[0]])
    assert np.allclose(A, np.outer(at, at), atol=1e-7)
    # rejected pairs: too far, normals too different, out of depth range, non-finite
    assert orc.icp_cost(v1, n1, v2 + [0, 0, 0.2], n1)[2] == 0
    assert orc.icp_cost(v1, n1, v2, np.array([[0.0, 0.6, -0.8]], np.float32))[2] == 0
    assert orc.icp_cost(v1 * [1, 1, 0.05], n1, v2 * [1, 1, 0.05], n1)[2] == 0
    assert orc.icp_cost(v1, n1, v2 * np.float32(np.inf), n1)[2] == 0


def _poses(n):
    out = []
    for k in range(n):
        M = np.eye(4)
        a = np.radians(0.3 * k)
        M[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        M[:3, 3] = [0.01 * k, 0.0, 0.005 * k]
        out.append(M.astype(np.float32))
    return out


def test_exact_tracker_recovers_the_synthetic_motion():
    synth = pkg().synth
    w, h = 320, 240
    fx, fy = synth.focal(w, h)
    poses = _poses(4)
    t = orc.OracleTracker(w, h, fx, fy, exact_jacobian=True)
    for k, P in enumerate(poses):
        d, _ = synth.make_frame(w, h, P, seed=k, invalid_frac=0.01, noise_mm=1)
        t.update(d)
        assert not t.lost
    got, want = t.pose(), np.linalg.inv(poses[0]) @ poses[-1]
    assert t.pairs > 0.5 * w * h * 0.5
    assert np.allclose(got[:3, :3], want[:3, :3], atol=4e-3)   # the motion is 1.6e-2 rad, 3.4e-2 m
    assert np.allclose(got[:3, 3], want[:3, 3], atol=4e-3), (got[:3, 3], want[:3, 3])


def test_reference_tracker_quirks():
    """Q18: position_ = vec3(vec4(position_, 1) * update) keeps a zero position at zero; the orientation moves."""
    synth = pkg().synth
    w, h = 160, 120
    fx, fy = synth.focal(w, h)
    t = orc.OracleTracker(w, h, fx, fy)
    for k, P in enumerate(_poses(3)):
        d, _ = synth.make_frame(w, h, P, seed=k, invalid_frac=0.01, noise_mm=1)
        t.update(d)
    assert np.array_equal(t.position(), np.zeros(3, np.float32))
    assert not np.array_equal(t.orientation(), np.eye(3, dtype=np.float32))
    assert np.allclose(t.pose()[:3, :3], t.orientation())


def test_blank_frames_lose_tracking():
    t = orc.OracleTracker(64, 48, 50.0, 50.0)
    z = np.zeros((48, 64), np.uint16)
    t.update(z)
    t.update(z)
    assert t.lost and t.pairs == 0
    assert np.array_equal(t.orientation(), np.eye(3, dtype=np.float32))  # no update was applied
