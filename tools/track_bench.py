"""Times sensor::RGBDCamera::update (osl_tracker_update) on resident VGA depth frames; with `ref` also the reference's
own RGBDCamera (oracle/_ref).  usage: python tools/track_bench.py [frames] [ref]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

P = graft.load_package()
lib = P.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
W, H = 640, 480
fx, fy = P.synth.focal(W, H)
frames = []
for k in range(8):
    pose = P.synth.orbit_pose(k)
    frames.append(torch.from_numpy(P.synth.make_frame(W, H, pose, seed=k)[0]).cuda())
cam = P.RGBDCamera(W, H, (fx, fy), exact_jacobian=True)
for k in range(3):
    cam.update(frames[k % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(3, 3 + n):
    lib.osl_tracker_update(cam._h, frames[k % 8].data_ptr(), None)
e1.record()
torch.cuda.synchronize()
print("tracker: %.3f ms / frame, lost=%s pairs=%d" % (e0.elapsed_time(e1) / n, cam.lost, cam.pairs))
if "ref" in sys.argv:
    from oracle import ref
    r = ref.RefTracker(W, H, fx, fy)
    host = [f.cpu().numpy() for f in frames]
    for k in range(3):
        r.update(host[k])
    ms = [r.update(host[k % 8]) for k in range(3, 13)]
    print("reference RGBDCamera::update: %.2f ms / frame" % (sum(ms) / len(ms)))
