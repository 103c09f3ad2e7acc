#!/usr/bin/env python
"""Phase checkpoints of k_structure / k_levels for one large integrate_voxels call (cfg2 scale)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
from tools.phase_profile import NAMES  # noqa: E402


def main():
    import torch
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    pkg = graft.load_package()
    lib = pkg.lib()
    V, T = pkg.synth.icosphere(4, 0.8)
    cen, col = pkg.meshToVoxelGrid(V, T, None, (0, 0, 0), 1.0, D)
    svo = pkg.SVO((0, 0, 0), 1.0, D, reserve_nodes=3 * cen.shape[0]).set_stage_timing(True)
    for _ in range(3):
        svo.integrate_voxels(cen, col)
        svo.sync()
    prof = (C.c_uint64 * 64)()
    lib.osl_debug_profile(prof, 64)
    print("voxels", cen.shape[0], "stage ms", svo.stage_times())
    for kern in ("k_structure", "k_levels"):
        for a, b, name in NAMES[kern]:
            print("  %-14s %-46s %9.1f us" % (kern, name, (prof[b] - prof[a]) / 1965.0))


if __name__ == "__main__":
    main()
