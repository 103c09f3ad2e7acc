#!/usr/bin/env python
"""Stage and phase times of svoFromVoxelGrid (osl_integrate_voxels) on a LARGE voxel grid: the cfg2 mesh voxelised at
depth D, in Morton order and shuffled.  Per-kernel CUDA-event times (osl_get_stage_times) and the SM-clock checkpoints
of CTA 0 (osl_debug_profile).  Run on the GPU box:  python tools/large_profile.py [depth] [reps]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

PHASES = [
    ("k_emit", [(0, 3, "whole (CTA 0)")]),
    ("k_structure", [(16, 17, "A analyze (all blocks of CTA 0)"), (17, 18, "wait for all flags"), (18, 19, "sums"),
                     (20, 21, "plan"), (21, 22, "C assign (all blocks of CTA 0)")]),
    ("k_levels", [(32, 35, "subtrees / leaves"), (35, 36, "wide levels"), (36, 37, "one-sided barrier"),
                  (37, 38, "stage narrow top"), (38, 39, "smem fold")]),
]


def main():
    import torch
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    pkg = graft.load_package()
    lib = pkg.lib()
    path = graft.asset("bunny_tex.obj")
    if path:
        V, T = pkg.synth.load_obj(path)
    else:
        V, T = pkg.synth.icosphere(4, 1.35)
    colors = np.random.default_rng(0).uniform(0.2, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    lo, hi = V.min(axis=0), V.max(axis=0)
    center = tuple(float(x) for x in (np.float32(0.5) * (lo + hi)))
    half = float(hi[0])
    cen, col = pkg.meshToVoxelGrid(V, T, colors, center, half, D)
    n = int(cen.shape[0])
    mhz = 1965.0
    for label in ("ordered", "shuffled"):
        if label == "shuffled":
            perm = torch.randperm(n, device="cuda")
            cen, col = cen[perm].contiguous(), col[perm].contiguous()
            del perm
        svo = pkg.SVO(center, half, D, reserve_nodes=max(1 << 20, int(2.7 * n))).set_stage_timing(True)
        for k in range(reps + 1):
            svo.integrate_voxels(cen, col)
            svo.sync()
            st = np.array(svo.stage_times()) * 1e3
            prof = (C.c_uint64 * 64)()
            lib.osl_debug_profile(prof, 64)
            cn = svo.counters()
            print("%s call %d: n=%d unique=%d split=%d nodes=%d | emit %.0f sort %.0f structure %.0f levels %.0f us (sum %.0f)"
                  % (label, k, n, cn.n_unique, cn.n_split, svo.size, st[0], st[1], st[2], st[3], st.sum()))
            if k in (0, reps):
                cp = (C.c_uint64 * 4096)()
                lib.osl_debug_cta_profile(cp)
                a = np.array(cp[:], dtype=np.float64).reshape(4, 1024)
                live = a[0] > 0
                if live.any():
                    q = lambda v: "min %.0f median %.0f max %.0f" % (v.min() / mhz, np.median(v) / mhz, v.max() / mhz)
                    print("      per CTA (%d): A took %s us | C took %s us" %
                          (int(live.sum()), q(a[1][live] - a[0][live]), q(a[3][live] - a[2][live])))
                    idx = np.nonzero(live)[0]
                    parts = np.array_split(idx[:-1], 12)
                    print("      by CTA index (12 groups): A " + " ".join("%.0f" % ((a[1][g] - a[0][g]).mean() / mhz) for g in parts) +
                          " | C " + " ".join("%.0f" % ((a[3][g] - a[2][g]).mean() / mhz) for g in parts))
                for kern, ph in PHASES:
                    for a, b, name in ph:
                        print("      %-12s %-36s %9.1f us" % (kern, name, (prof[b] - prof[a]) / mhz))
        svo.close()


if __name__ == "__main__":
    main()
