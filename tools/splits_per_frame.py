import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
P = g.load_package()
import torch
D, w, h = 16, 640, 480
center, half = P.synth.tree_params(D)
fx, fy = P.synth.focal(w, h)
svo = P.SVO(center, half, D, reserve_nodes=1 << 24)
S = []
for k in range(120):
    pose = P.synth.orbit_pose(k)
    d, c = P.synth.make_frame(w, h, pose, seed=k)
    svo.integrate_depth(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), fx, fy, pose)
    svo.sync()
    S.append(svo.counters().n_split)
print("splits per frame:", S)
print("frames with zero splits: %d of %d" % (sum(1 for s in S[20:] if s == 0), len(S) - 20))
