#!/usr/bin/env python
"""Phase-level timing of the integrate kernels on the bench workload (640x480, depth 16), from the SM-clock
checkpoints CTA 0 writes (osl_debug_profile) plus the per-kernel CUDA-event times (osl_get_stage_times).
Run on the GPU box:  python tools/phase_profile.py [frames]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

NAMES = {
    "k_emit": [(0, 1, "load+backproject+keys"), (1, 2, "hash de-dup"), (2, 3, "append")],
    "k_sort_bucket": [(8, 9, "scan list"), (9, 10, "smem radix sort"), (10, 11, "write")],
    "k_structure": [(16, 17, "A analyze (prefix, run-min, walk)"), (17, 18, "wait for all flags / barrier"),
                    (18, 19, "sum vectors / column scans"), (19, 20, "(barrier)"), (20, 21, "B2 plan"),
                    (21, 22, "C assign (short walk, level lists, tile init)"),
                    (16, 24, "  A: prologue (n, cur, splitters)"), (24, 25, "  A: key + predecessor + run-min"),
                    (25, 26, "  A: walk (thread 0)"), (26, 17, "  A: counts, block sync, publish"),
                    (21, 27, "  C: pass 1 (ballots)"), (27, 28, "  C: warp scan"), (28, 22, "  C: pass 2 (walk, lists, tiles)")],
    "k_levels": [(32, 35, "leaves"), (35, 36, "wide levels (grid barrier each)"), (36, 37, "one-sided barrier"),
                 (37, 38, "stage narrow top"), (38, 39, "smem fold")],
}


def main():
    import torch
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
    D = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    pkg = graft.load_package()
    lib = pkg.lib()
    center, half = pkg.synth.tree_params(D)
    fx, fy = pkg.synth.focal(w, h)
    svo = pkg.SVO(center, half, D, reserve_nodes=1 << 24).set_stage_timing(True)
    mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
    acc = {}
    stage = np.zeros(4)
    cnt = 0
    for k in range(frames):
        pose = pkg.synth.orbit_pose(k)
        depth, rgb = pkg.synth.make_frame(w, h, pose, seed=k)
        d, c = torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda()
        torch.cuda.synchronize()
        svo.integrate_depth(d, c, fx, fy, pose)
        svo.sync()
        if k < frames // 2:
            continue
        prof = (C.c_uint64 * 64)()
        lib.osl_debug_profile(prof, 64)
        for kern, phases in NAMES.items():
            for a, b, name in phases:
                acc[(kern, name)] = acc.get((kern, name), 0.0) + (prof[b] - prof[a]) / float(mhz)
        stage += np.array(svo.stage_times()) * 1e3
        cnt += 1
    cn = svo.counters()
    print("workload %dx%d D=%d: N=%d V=%d U=%d S=%d  (SM clock %d MHz, %d frames averaged)" %
          (w, h, D, cn.n_points, cn.n_valid, cn.n_unique, cn.n_split, mhz, cnt))
    for i, kern in enumerate(["k_emit", "k_sort_bucket", "k_structure", "k_levels"]):
        print("%-14s event %.1f us" % (kern, stage[i] / cnt))
        for a, b, name in NAMES[kern]:
            print("    %-36s %7.2f us" % (name, acc[(kern, name)] / cnt))


if __name__ == "__main__":
    main()
