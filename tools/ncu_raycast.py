#!/usr/bin/env python
"""The bench map (25 pipelined orbit frames, depth 16) and three 640x480 raycasts from the last pose -- a short command
for `ncu --set full -k regex:k_raycast` (profiles/r02_ncu_full_summary_raycast.csv)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    import torch
    pkg = graft.load_package()
    W, H, D = 640, 480, 16
    center, half = pkg.synth.tree_params(D)
    fx, fy = pkg.synth.focal(W, H)
    svo = pkg.SVO(center, half, D, reserve_nodes=1 << 24)
    pose = None
    for k in range(25):
        pose = pkg.synth.orbit_pose(k)
        d, c = pkg.synth.make_frame(W, H, pose, seed=k)
        svo.integrate_depth(d, c, fx, fy, pose)
    svo.sync()
    view = (np.diag([-1.0, 1.0, -1.0, 1.0]) @ np.linalg.inv(pose.astype(np.float64))).astype(np.float32)
    out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        svo.raycast_device(out, W, H, 45.0, view)
    torch.cuda.synchronize()
    print("nodes", svo.size)


if __name__ == "__main__":
    main()
