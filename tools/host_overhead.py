#!/usr/bin/env python
"""Host enqueue cost per frame vs device time per frame (pipelined, resident inputs), bench workload."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    import torch
    pkg = graft.load_package()
    lib = pkg.lib()
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 480
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    R = 32
    center, half = pkg.synth.tree_params(D)
    fx, fy = pkg.synth.focal(W, H)
    frames = [pkg.synth.make_frame(W, H, pkg.synth.orbit_pose(k), seed=k) for k in range(R)]
    dd = [torch.from_numpy(f[0]).cuda() for f in frames]
    cc = [torch.from_numpy(f[1]).cuda() for f in frames]
    poses = [pkg.capi._f(pkg.capi.mat_colmajor(pkg.synth.orbit_pose(k))) for k in range(R)]
    torch.cuda.synchronize()
    for piped in (True, False):
        svo = pkg.SVO(center, half, D, reserve_nodes=1 << 24).set_pipeline(piped)
        for k in range(40):
            lib.osl_integrate_depth(svo._h, dd[k % R].data_ptr(), cc[k % R].data_ptr(), W, H, fx, fy, poses[k % R], None)
        svo.sync()
        K = 400
        t0 = time.perf_counter()
        for k in range(K):
            lib.osl_integrate_depth(svo._h, dd[k % R].data_ptr(), cc[k % R].data_ptr(), W, H, fx, fy, poses[k % R], None)
        t1 = time.perf_counter()
        svo.sync()
        t2 = time.perf_counter()
        print("pipelined=%d: host enqueue %.1f us/frame, total %.1f us/frame" %
              (piped, (t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6))
        # un-throttled host cost: 3 calls after a full sync (the host never waits for the device here), repeated
        acc = 0.0
        for rep in range(50):
            svo.sync()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(3):
                lib.osl_integrate_depth(svo._h, dd[k % R].data_ptr(), cc[k % R].data_ptr(), W, H, fx, fy, poses[k % R], None)
            acc += time.perf_counter() - t0
        print("   un-throttled host cost %.1f us/call" % (acc / 150 * 1e6))
        svo.close()


if __name__ == "__main__":
    main()
