import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
P = g.load_package()
import torch
V, T = P.synth.load_obj(g.asset("bunny_tex.obj"))
colors = np.random.default_rng(0).uniform(0.2, 1.0, size=(T.shape[0], 4)).astype(np.float32)
lo, hi = V.min(axis=0), V.max(axis=0)
center = tuple(float(x) for x in (np.float32(0.5) * (lo + hi))); half = float(hi[0])
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cen, col = P.meshToVoxelGrid(V, T, colors, center, half, 12)
    torch.cuda.synchronize(); print("call %d: %.1f ms, %d voxels" % (k, (time.perf_counter() - t0) * 1e3, cen.shape[0]))
    del cen, col
