"""Deterministic synthetic RGB-D input (scene S0 "sphere-in-room"): a sphere r = 0.5 m at
(0, 0, 1.5) inside an axis-aligned room x in [-2, 2], y in [-1.5, 1.5], z in [-2.5, 2.5]; pin-hole camera with the
reference's Kinect focal lengths (cone_tracing_kernels.cu:45-46) scaled to the resolution.  Depth is uint16
millimetres with 2 % invalid pixels and +-2 mm noise, colour a 10 cm 3-D checker XOR a per-pixel hash.
Pure numpy (host side): this only produces INPUT frames, it is not part of the hot path."""
import math

import numpy as np

ROOM_LO = np.array([-2.0, -1.5, -2.5])
ROOM_HI = np.array([2.0, 1.5, 2.5])
SPHERE_C = np.array([0.0, 0.0, 1.5])
SPHERE_R = 0.5


def focal(w, h):
    return np.float32(532.57 * w / 640.0), np.float32(531.54 * h / 480.0)


def orbit_pose(k, radius=1.2, step_deg=0.36):
    """Camera-to-world pose k of the cfg3 orbit: circle of `radius` around the sphere centre at height 0,
    looking at the centre.  Returns a 4x4 float32 matrix applied to camera-space points (main.cpp:40)."""
    a = math.radians(step_deg * k)
    eye = SPHERE_C + radius * np.array([math.sin(a), 0.0, -math.cos(a)])
    fwd = SPHERE_C - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(up, fwd)
    right /= np.linalg.norm(right)
    up2 = np.cross(fwd, right)
    M = np.eye(4)
    M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = right, up2, fwd, eye
    return M.astype(np.float32)


def _raycast_scene(origin, dirs):
    """distance t along unit dirs to the nearest of {sphere, room walls}; returns t and hit points"""
    oc = origin - SPHERE_C
    b = dirs @ oc
    c = oc @ oc - SPHERE_R ** 2
    disc = b * b - c
    t_s = np.where(disc > 0, -b - np.sqrt(np.maximum(disc, 0)), np.inf)
    t_s = np.where(t_s > 1e-6, t_s, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_lo = (ROOM_LO - origin) / dirs
        t_hi = (ROOM_HI - origin) / dirs
    t_wall = np.where(dirs > 0, t_hi, t_lo)
    t_wall = np.where(np.isfinite(t_wall) & (t_wall > 0), t_wall, np.inf).min(axis=1)
    t = np.minimum(t_s, t_wall)
    return t, origin + dirs * t[:, None]


def make_frame(w=640, h=480, pose=None, seed=0, invalid_frac=0.02, noise_mm=2):
    """Returns depth (h,w) uint16 mm, rgb (h,w,3) uint8 for camera pose `pose` (camera-to-world, 4x4)."""
    pose = np.eye(4) if pose is None else np.asarray(pose, dtype=np.float64)
    fx, fy = focal(w, h)
    xs = np.arange(w) - w / 2.0
    ys = h / 2.0 - np.arange(h)
    X, Y = np.meshgrid(xs / float(fx), ys / float(fy))
    dirs_c = np.stack([X.ravel(), Y.ravel(), np.ones(w * h)], axis=1)
    zscale = 1.0 / np.linalg.norm(dirs_c, axis=1)  # z of the unit direction
    dirs_c *= zscale[:, None]
    R, tr = pose[:3, :3], pose[:3, 3]
    t, hit = _raycast_scene(tr, dirs_c @ R.T)
    z_mm = np.where(np.isfinite(t), t * zscale * 1000.0, 0.0)
    rng_inv = np.random.default_rng(7 + seed)
    rng_noise = np.random.default_rng(11 + seed)
    noise = rng_noise.integers(-noise_mm, noise_mm + 1, size=w * h)
    depth = np.clip(np.rint(z_mm) + noise, 0, 15000)
    depth = np.where(z_mm > 0, depth, 0)
    depth = np.where(rng_inv.random(w * h) < invalid_frac, 0, depth).astype(np.uint16)
    cell = np.floor(hit * 10.0).astype(np.int64)
    checker = ((cell[:, 0] + cell[:, 1] + cell[:, 2]) & 1).astype(np.uint8)
    base = np.stack([60 + 150 * checker, 200 - 120 * checker, 90 + 60 * ((cell[:, 0] & 3) == 0)], axis=1).astype(np.uint8)
    pix = np.arange(w * h, dtype=np.uint64) + np.uint64(13 + seed)
    hsh = (pix * np.uint64(2654435761)) >> np.uint64(7)
    noise_rgb = np.stack([hsh & np.uint64(31), (hsh >> np.uint64(5)) & np.uint64(31),
                          (hsh >> np.uint64(10)) & np.uint64(31)], axis=1).astype(np.uint8)
    rgb = base ^ noise_rgb
    return depth.reshape(h, w), rgb.reshape(h, w, 3)


def tree_params(max_depth, resolution=0.01):
    """half_edge = resolution * 2^D built by repeated doubling in float32, centre 0."""
    he = np.float32(resolution)
    for _ in range(max_depth):
        he = np.float32(he * np.float32(2.0))
    return (0.0, 0.0, 0.0), float(he)
