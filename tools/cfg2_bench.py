#!/usr/bin/env python
"""BASELINE configs[1] shape at full size: a ~5 k-triangle mesh voxelised to a depth-12 SVO, rendered at 1920x1080.
(The reference's bunny_tex.obj lives in /root/reference, which the GPU box does not have: a 5 120-triangle
icosphere of the same extent stands in.)  Prints one JSON line.  python tools/cfg2_bench.py [depth] [subdiv] [reps]"""
import json
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
    subdiv = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    pkg = graft.load_package()
    V, T = pkg.synth.icosphere(subdiv, 0.8)
    rng = np.random.default_rng(0)
    colors = rng.uniform(0.2, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    center, half = (0.0, 0.0, 0.0), 1.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cen, col = pkg.meshToVoxelGrid(V, T, colors, center, half, D)
    torch.cuda.synchronize()
    t_vox = time.perf_counter() - t0
    t0 = time.perf_counter()
    cen, col = pkg.meshToVoxelGrid(V, T, colors, center, half, D)
    torch.cuda.synchronize()
    t_vox2 = time.perf_counter() - t0
    n = cen.shape[0]
    if len(sys.argv) > 4 and sys.argv[4] == "shuffle":  # a grid that does NOT arrive in Morton order: the radix sort runs
        perm = torch.randperm(n, device="cuda")
        cen, col = cen[perm].contiguous(), col[perm].contiguous()
    svo = pkg.SVO(center, half, D, reserve_nodes=max(1 << 20, 3 * n))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for k in range(reps):
        svo.integrate_voxels(cen, col)
        ev[k + 1].record()
    svo.sync()
    torch.cuda.synchronize()
    ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(reps)]
    svo.set_stage_timing(True)
    svo.integrate_voxels(cen, col)
    stage = svo.stage_times()
    svo.set_stage_timing(False)
    cn = svo.counters()
    first_bytes = None
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # steady state (re-observation of the same grid: no splits): bytes of the LAST call
    steady_ms = float(np.median(ms[reps // 2:])) if reps > 2 else ms[-1]
    steady_gbs = cn.algorithmic_bytes / (steady_ms / 1e3) / 1e9
    W, H = 1920, 1080
    view = np.eye(4, dtype=np.float32)
    view[2, 3] = -2.5
    out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        svo.raycast_device(out, W, H, 45.0, view)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(5):
        svo.raycast_device(out, W, H, 45.0, view)
    r1.record()
    torch.cuda.synchronize()
    ray_ms = r0.elapsed_time(r1) / 5
    st = pkg.RaycastStats()
    img = svo.raycast(W, H, 45.0, view, stats=st)
    print(json.dumps({
        "workload": "cfg2 shape: %d-triangle mesh -> depth-%d SVO (%d voxels, %d nodes), %dx%d raycast" %
                    (T.shape[0], D, n, svo.size, W, H),
        "voxelise_ms": {"first_call": t_vox * 1e3, "second_call": t_vox2 * 1e3},
        "integrate_voxels_ms": {"first": ms[0], "steady_median": steady_ms,
                                "stages": {"k_emit": stage[0], "k_sort": stage[1], "k_structure": stage[2],
                                           "k_levels": stage[3]}},
        "integrate_steady": {"algorithmic_bytes": int(cn.algorithmic_bytes), "GBps": steady_gbs,
                             "frac_of_measured_hbm_peak": steady_gbs / peak, "U": int(cn.n_unique)},
        "raycast": {"ms": ray_ms, "mrays_per_s": W * H / (ray_ms / 1e3) / 1e6, "steps_per_ray": st.steps / float(st.rays),
                    "lit_pixels": int(np.count_nonzero(img[..., :3].sum(axis=2)))},
    }))


if __name__ == "__main__":
    main()
