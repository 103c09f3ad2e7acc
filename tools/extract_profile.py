#!/usr/bin/env python
"""Time of extractVoxelGridFromSVO (osl_extract_voxels) on the cfg2 tree (bunny, depth 12, 57.5 M voxels).
Run on the GPU box:  python tools/extract_profile.py [depth]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    import torch
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    P = graft.load_package()
    lib = P.lib()
    path = graft.asset("bunny_tex.obj")
    V, T = P.synth.load_obj(path) if path else P.synth.icosphere(4, 1.35)
    colors = np.random.default_rng(0).uniform(0.2, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    lo, hi = V.min(axis=0), V.max(axis=0)
    center = tuple(float(x) for x in (np.float32(0.5) * (lo + hi)))
    half = float(hi[0])
    cen, col = P.meshToVoxelGrid(V, T, colors, center, half, D)
    n = int(cen.shape[0])
    svo = P.SVO(center, half, D, reserve_nodes=int(3.8 * n))
    for _ in range(3):
        svo.integrate_voxels(cen, col)
    svo.sync()
    cnt = C.c_int64()
    for k in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        P.capi._check(lib.osl_extract_voxels(svo._h, D, None, None, None, 0, C.byref(cnt), None), "count")
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        c = torch.empty((cnt.value, 4), dtype=torch.float32, device="cuda")
        k4 = torch.empty((cnt.value, 4), dtype=torch.float32, device="cuda")
        keys = torch.empty((cnt.value,), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        P.capi._check(lib.osl_extract_voxels(svo._h, D, c.data_ptr(), k4.data_ptr(), keys.data_ptr(), cnt.value,
                                             C.byref(cnt), None), "extract")
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print("call %d: %d voxels of %d; count %.2f ms, extract %.2f ms" % (k, cnt.value, n, (t1 - t0) * 1e3, (t3 - t2) * 1e3))


if __name__ == "__main__":
    main()
