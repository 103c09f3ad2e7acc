#!/usr/bin/env python
"""e2e (pinned host frames) rate of the bench workload through the raw C ABI, minimal Python per call.
    python tools/e2e_probe.py [frames]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    import torch
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    pkg = graft.load_package()
    lib = pkg.lib()
    W, H, D, RING = 640, 480, 16, 136
    center, half = pkg.synth.tree_params(D)
    fx, fy = pkg.synth.focal(W, H)
    depths, rgbs, poses = [], [], []
    for k in range(RING):
        pose = pkg.synth.orbit_pose(k)
        d, c = pkg.synth.make_frame(W, H, pose, seed=k)
        depths.append(d); rgbs.append(c); poses.append(pkg.capi._f(pkg.capi.mat_colmajor(pose)))
    hd = torch.from_numpy(np.stack(depths)).pin_memory()
    hc = torch.from_numpy(np.stack(rgbs)).pin_memory()
    ptr = [(hd[j].data_ptr(), hc[j].data_ptr()) for j in range(RING)]
    svo = pkg.SVO(center, half, D, reserve_nodes=1 << 24)
    f = lib.osl_integrate_depth_host
    for k in range(40):
        f(svo._h, ptr[k % RING][0], ptr[k % RING][1], W, H, fx, fy, poses[k % RING], None)
    svo.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for k in range(40, 40 + K):
        f(svo._h, ptr[k % RING][0], ptr[k % RING][1], W, H, fx, fy, poses[k % RING], None)
    t_host = time.perf_counter() - t0
    svo.join(None)
    e1.record()
    svo.sync()
    us = e0.elapsed_time(e1) * 1e3 / K
    print("%s: %.2f us/frame = %.0f frames/s on the device; host spent %.2f us/frame enqueuing" %
          (os.environ.get("OSL_FZ_HOST_EVENT") and "event edge" or (os.environ.get("OSL_NO_FUSED") and "four kernels" or "flag"),
           us, 1e6 / us, t_host * 1e6 / K))


if __name__ == "__main__":
    main()
