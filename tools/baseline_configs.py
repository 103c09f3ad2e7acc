#!/usr/bin/env python
"""the reference path (its own kernels re-targeted to sm_100a + its host sequence, 1 host
thread) and this library, side by side, on the BASELINE configurations that are not the bench headline.
  cfg1  one 640x480 frame re-inserted into a depth-8 SVO (unmodified reference, D <= 10) + 640x480 raycast
  cfg3  640x480 orbit, incremental fusion into a depth-14 SVO (ref + 64-bit patch), first 200 frames
  cfg4  1280x960 frames into a depth-16 SVO (ref + 64-bit patch), 60 frames
Prints one JSON line per configuration.  Needs oracle/_ref (built where /root/reference is mounted)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from oracle import ref as R  # noqa: E402


def run(name, W, H, D, frames, same_frame, patched, ray=False):
    import torch
    pkg = graft.load_package()
    center, half = pkg.synth.tree_params(D)
    fx, fy = pkg.synth.focal(W, H)
    ring = 1 if same_frame else min(frames, 48)
    data = []
    for k in range(ring):
        pose = pkg.synth.orbit_pose(k)
        d, c = pkg.synth.make_frame(W, H, pose, seed=k)
        data.append((torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), pose))
    torch.cuda.synchronize()
    warm = 10
    # reference
    t = R.RefSVO(center, half, D, patched64=patched)
    for k in range(warm):
        d, c, pose = data[k % ring]
        t.integrate_depth_dev(d.data_ptr(), c.data_ptr(), W, H, fx, fy, pose)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(warm, warm + frames):
        d, c, pose = data[k % ring]
        t.integrate_depth_dev(d.data_ptr(), c.data_ptr(), W, H, fx, fy, pose)
    torch.cuda.synchronize()
    ref_fps = frames / (time.perf_counter() - t0)
    out = {"config": name, "frames": frames, "reference_frames_per_s": ref_fps, "reference_nodes": t.size}
    if ray:  # row (f)-1: SVO -> voxel list (svo.cu:699-745)
        t.extract_voxels(D)
        t0 = time.perf_counter()
        for _ in range(5):
            rc, rk = t.extract_voxels(D)
        out["reference_extract_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        out["reference_extract_voxels"] = int(rc.shape[0])
    view = (np.diag([-1.0, 1.0, -1.0, 1.0]) @ np.linalg.inv(np.asarray(data[0][2], dtype=np.float64))).astype(np.float32)
    if ray:
        t.raycast(W, H, 45.0, view, want_image=False)
        _, ms = t.raycast(W, H, 45.0, view, want_image=False)
        out["reference_raycast_mrays_per_s"] = W * H / (ms / 1e3) / 1e6
    del t
    # ours (pipelined, resident frames)
    s = pkg.SVO(center, half, D, reserve_nodes=1 << 22).set_pipeline(True)
    lib = pkg.lib()
    poses = [pkg.capi._f(pkg.capi.mat_colmajor(p)) for _, _, p in data]
    for k in range(warm):
        d, c, _ = data[k % ring]
        lib.osl_integrate_depth(s._h, d.data_ptr(), c.data_ptr(), W, H, fx, fy, poses[k % ring], None)
    s.sync()
    t0 = time.perf_counter()
    for k in range(warm, warm + frames):
        d, c, _ = data[k % ring]
        lib.osl_integrate_depth(s._h, d.data_ptr(), c.data_ptr(), W, H, fx, fy, poses[k % ring], None)
    s.sync()
    out["ours_frames_per_s"] = frames / (time.perf_counter() - t0)
    out["ours_nodes"] = s.size
    if ray:
        s.extract_voxels(D)
        t0 = time.perf_counter()
        for _ in range(5):
            c_, k_, keys_ = s.extract_voxels(D)
        out["ours_extract_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        out["ours_extract_voxels"] = int(c_.shape[0])
        img = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            s.raycast_device(img, W, H, 45.0, view)
        e0.record()
        for _ in range(10):
            s.raycast_device(img, W, H, 45.0, view)
        e1.record()
        torch.cuda.synchronize()
        out["ours_raycast_mrays_per_s"] = W * H / (e0.elapsed_time(e1) / 10 / 1e3) / 1e6
    out["speedup_integrate"] = out["ours_frames_per_s"] / ref_fps
    print(json.dumps(out))


if __name__ == "__main__":
    run("cfg1: 640x480 frame re-inserted, depth-8 SVO, unmodified reference", 640, 480, 8, 200, True, False, ray=True)
    run("cfg3: 640x480 orbit, incremental, depth-14 SVO, ref+64-bit patch", 640, 480, 14, 200, False, True)
    run("cfg4 frame size: 1280x960 orbit, depth-16 SVO, ref+64-bit patch", 1280, 960, 16, 60, False, True)
