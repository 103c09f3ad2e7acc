#!/usr/bin/env python
"""Timeline of the k_frame pipeline on the bench workload (640x480 -> depth 16, resident frames): for a run of
consecutive launches, when each role's first CTA started and its last CTA ended (%globaltimer, osl_debug_trace).
    python tools/frame_timeline.py [frames]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as graft  # noqa: E402

ROLES = ["structure", "values", "emit", "sort", "arrival"]


def main():
    import torch
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    host = len(sys.argv) > 2 and sys.argv[2] == "host"  # pinned host frames through osl_integrate_depth_host
    w, h, D = 640, 480, 16
    pkg = graft.load_package()
    lib = pkg.lib()
    center, half = pkg.synth.tree_params(D)
    fx, fy = pkg.synth.focal(w, h)
    svo = pkg.SVO(center, half, D, reserve_nodes=1 << 24).set_pipeline(True)
    keep = []
    for k in range(40 + frames):
        pose = pkg.synth.orbit_pose(k)
        depth, rgb = pkg.synth.make_frame(w, h, pose, seed=k)
        if host:
            keep.append((torch.from_numpy(depth).pin_memory().numpy(), torch.from_numpy(rgb).pin_memory().numpy(), pose))
        else:
            keep.append((torch.from_numpy(depth).cuda(), torch.from_numpy(rgb).cuda(), pose))
    torch.cuda.synchronize()
    integrate = svo.integrate_depth_host if host else svo.integrate_depth
    for d, c, pose in keep[:40]:
        integrate(d, c, fx, fy, pose)
    svo.sync()
    lib.osl_debug_trace(svo._h, 1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for d, c, pose in keep[40:]:
        integrate(d, c, fx, fy, pose)
    svo.join(None)
    e1.record()
    svo.sync()
    out = (C.c_uint64 * (32 * 5 * 2))()
    lib.osl_debug_trace(svo._h, 0, out)
    a = np.array(out, dtype=np.uint64).reshape(32, 5, 2).astype(np.float64)
    n_launch = min(32, frames + 3)
    t0 = a[:n_launch, :, 0][a[:n_launch, :, 0] < 1e19].min()
    print("%d frames, %.2f us per frame by CUDA events; times in us since the first CTA" % (frames, e0.elapsed_time(e1) * 1e3 / frames))
    print("launch  " + "  ".join("%-17s" % r for r in ROLES) + "  launch span   start-to-start")
    prev = None
    for i in range(n_launch):
        cells = []
        lo, hi = 1e30, 0.0
        for r in range(5):
            s, e = a[i, r]
            if s > 1e19:
                cells.append("%-17s" % "-")
                continue
            cells.append("%7.1f - %7.1f" % ((s - t0) / 1e3, (e - t0) / 1e3))
            if r < 4:
                lo, hi = min(lo, s), max(hi, e)
        gap = "" if prev is None else "%6.1f" % ((lo - prev) / 1e3)
        print("%4d    %s  %6.1f        %s" % (i, "  ".join(cells), (hi - lo) / 1e3, gap))
        prev = lo
    # phases of CTA 0 of each role in the LAST launch that ran it (SM-clock checkpoints, osl_debug_profile)
    from phase_profile import NAMES
    prof = (C.c_uint64 * 128)()
    lib.osl_debug_profile(prof, 128)
    mhz = 1965.0
    # per-CTA phase lengths of the LAST structure launch (SM clocks: only differences inside one CTA are meaningful)
    cp = (C.c_uint64 * 4096)()
    lib.osl_debug_cta_profile(cp)
    q = np.array(cp[:], dtype=np.float64).reshape(4, 1024)
    live = (q[0] > 0) & (q[3] > q[2]) & (q[1] > q[0])
    if live.any():
        idx = np.nonzero(live)[0]
        print("structure role, last launch, per CTA: analyze / exchange+plan / assign [us]")
        print("   " + "  ".join("%d:%.1f/%.1f/%.1f" % (i, (q[1][i] - q[0][i]) / mhz, (q[2][i] - q[1][i]) / mhz, (q[3][i] - q[2][i]) / mhz)
                              for i in idx))
    for kern, phases in NAMES.items():
        print(kern)
        for a_, b_, name in phases:
            print("    %-44s %7.2f us" % (name, (prof[b_] - prof[a_]) / mhz))
    print("warp 0 of CTA 0: solo prefix %.2f us, warp-wide loop from level %d" % ((prof[96] - prof[28]) / mhz, prof[97]))
    print("assign pass 2 of CTA 0, per warp: us after the pass began / first level d0 / any split")
    print("   " + "  ".join("%.1f/%d/%d" % ((prof[64 + w] - prof[28]) / mhz, prof[80 + w] & 0xFF, prof[80 + w] >> 8)
                           for w in range(16)))


if __name__ == "__main__":
    main()
