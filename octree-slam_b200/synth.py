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


def icosphere(subdiv=3, radius=0.8, center=(0.0, 0.0, 0.0)):
    """Triangle mesh of a sphere (20 * 4^subdiv triangles) -- procedural input for the mesh voxeliser (the GPU box has
    no access to the reference's objs/).  Returns (vertices float32 [nv,3], triangles int32 [nt,3])."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    V = (np.array(v) * radius + np.asarray(center)).astype(np.float32)
    return V, np.array(f, dtype=np.int32)


def load_obj(path, with_uv=False):
    """Minimal Wavefront OBJ reader (v / vt / f, polygons fan-triangulated like the reference's objUtil,
    obj.cpp:44-90).  with_uv: also returns float32 [nt, 2], the texture coordinate of each triangle's FIRST corner
    (what the reference's ColorShader samples, voxelization.cu:113-126), or None when the file has none."""
    vs, vts, fs, fuv = [], [], [], []
    with open(path) as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(x) for x in p[1:4]])
            elif p[0] == "vt":
                vts.append([float(x) for x in p[1:3]])
            elif p[0] == "f":
                toks = [tok.split("/") for tok in p[1:]]
                idx = [int(t[0]) for t in toks]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                uv = [int(t[1]) if len(t) > 1 and t[1] else 0 for t in toks]
                uv = [i - 1 if i > 0 else (len(vts) + i if i < 0 else -1) for i in uv]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
                    fuv.append(uv[0])
    V, T = np.array(vs, dtype=np.float32), np.array(fs, dtype=np.int32)
    if not with_uv:
        return V, T
    if not vts:
        return V, T, None
    vt = np.array(vts, dtype=np.float32)
    fu = np.array(fuv, dtype=np.int64)
    uv = np.where((fu >= 0)[:, None], vt[np.clip(fu, 0, len(vts) - 1)], 0.0).astype(np.float32)
    return V, T, uv


def load_bmp(path):
    """24-bit BMP as the reference reads it (scene.cpp:38-62: 54-byte header, BGR bytes, rows as stored, no padding
    handling) -> float32 [height, width, 3] RGB in [0, 1]."""
    raw = open(path, "rb").read()
    w = int.from_bytes(raw[18:22], "little", signed=True)
    h = int.from_bytes(raw[22:26], "little", signed=True)
    data = np.frombuffer(raw, dtype=np.uint8, count=3 * w * h, offset=54).reshape(h, w, 3)
    return (data[..., ::-1].astype(np.float32) / np.float32(255.0))


def triangle_colors(uv, tex, wrap=False):
    """The reference's per-triangle flat colour (ColorShader, voxelization.cu:90-139 + createVoxelGrid :219-236): the
    texel at the first corner's (u, v) -- int(u * width), int(v * height), clamped here where the reference indexes
    unchecked -- quantised to 8 bits and returned as byte / 255; no texture coordinates: texel 0; no texture: green.
    -> float32 [nt, 4] (alpha 0, as the reference leaves it)."""
    nt = uv.shape[0] if uv is not None else 0
    out = np.zeros((nt, 4), dtype=np.float32)
    if tex is None:
        out[:, 1] = 1.0
        return out
    h, w = tex.shape[:2]
    if wrap:  # tiling texture coordinates (sponza): the fractional part, as a renderer would sample them
        uv = uv - np.floor(uv)
    tx = np.clip((uv[:, 0] * w).astype(np.int64), 0, w - 1)
    ty = np.clip((uv[:, 1] * h).astype(np.int64), 0, h - 1)
    c = np.clip(tex[ty, tx], 0.0, 1.0)
    out[:, :3] = ((c * np.float32(255.0)).astype(np.int64) & 0xFF) / 255.0
    return out
