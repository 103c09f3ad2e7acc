#!/usr/bin/env python
"""BASELINE configs 2 and 5 on their NAMED inputs (staged from the reference tree into baseline/_assets by
__graft_entry__.build()):

  cfg2  objs/bunny_tex.obj (4 968 triangles, textures/texture1.bmp) voxelised into a depth-12 SVO, 1920x1080 raycast
  cfg5  objs/crytek-sponza/sponza.obj (262 267 triangles after fan triangulation) voxelised as deep as the node
        layout allows, 3840x2160 raycast in interleaved row bands over all ranks

    python tools/cfg_mesh_bench.py cfg2|cfg5 [--depth D] [--reps R]
    python -m torch.distributed.run --nproc-per-node N ... tools/cfg_mesh_bench.py cfg5

The tree cube follows Scene::voxelizeMeshes (scene.cpp:64-85): centre = bounding-box mid-point, half edge =
bbox.bbox1.x.  Prints one JSON line (rank 0).  `run()` is what bench.py calls for its `cfg2` / `cfg5` objects."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

CONFIGS = {
    # name: (asset, default depth, render size, scale applied to the vertices, why)
    "cfg2": ("bunny_tex.obj", 12, (1920, 1080), 1.0),
    # sponza is modelled in centimetre-like units (extent 3 700); the raycaster's max_range is 10 world units
    # (cone_tracing_kernels.cu:24), so the mesh is brought to metres-like units first (x 1/400 -> half edge 4.5)
    "cfg5": ("sponza.obj", 12, (3840, 2160), 1.0 / 400.0),
}


def look_at(eye, target, up=(0.0, 1.0, 0.0)):
    """view matrix (world -> camera) of a camera at `eye` looking at `target` down its -z axis (the reference renderer's
    convention, cone_tracing_kernels.cu:161-167)"""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    M = np.eye(4)
    M[0, :3], M[1, :3], M[2, :3] = s, u, -f
    M[:3, 3] = -M[:3, :3] @ eye
    return M.astype(np.float32)


def run(name, depth=None, reps=3, rank=0, world=1, device=0, verbose=False):
    import torch
    pkg = graft.load_package()
    asset_name, d_default, (W, H), scale = CONFIGS[name]
    D = int(depth or d_default)
    path = graft.asset(asset_name)
    stand_in = path is None
    if stand_in:  # the GPU box without staged assets: say so, use a mesh of the same size class
        V, T = pkg.synth.icosphere(4 if name == "cfg2" else 7, 1.35)
        colors = np.random.default_rng(0).uniform(0.2, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    else:
        V, T, uv = pkg.synth.load_obj(path, with_uv=True)
        V = (V * np.float32(scale)).astype(np.float32)
        tex_path = graft.asset("texture1.bmp")
        tex = pkg.synth.load_bmp(tex_path) if tex_path else None
        colors = pkg.synth.triangle_colors(uv if uv is not None else np.zeros((T.shape[0], 2), np.float32), tex,
                                           wrap=(name == "cfg5"))
    lo, hi = V.min(axis=0), V.max(axis=0)
    center = tuple(float(x) for x in (np.float32(0.5) * (lo + hi)))
    half = float(hi[0])  # scene.cpp:78: bbox.bbox1.x (quirk Q10), not an extent
    torch.cuda.set_device(device)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def sync():
        torch.cuda.synchronize()

    # ---- mesh -> VoxelGrid (sparse voxeliser, Morton order)
    sync()
    t0 = time.perf_counter()
    cen, col = pkg.meshToVoxelGrid(V, T, colors, center, half, D, device=device)
    sync()
    vox_first = time.perf_counter() - t0
    del cen, col
    t0 = time.perf_counter()
    cen, col = pkg.meshToVoxelGrid(V, T, colors, center, half, D, device=device)
    sync()
    vox_ms = (time.perf_counter() - t0) * 1e3
    n = int(cen.shape[0])
    # ---- VoxelGrid -> SVO (svoFromVoxelGrid): the first call builds the tree, later calls re-observe it
    svo = pkg.SVO(center, half, D, reserve_nodes=max(1 << 20, int(2.7 * n)), device=device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 2)]
    ev[0].record()
    for k in range(reps + 1):
        svo.integrate_voxels(cen, col)
        ev[k + 1].record()
    svo.sync()
    sync()
    ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(reps + 1)]
    cn = svo.counters()
    nodes = svo.size
    steady_ms = float(np.median(ms[1:]))
    # the same grid in a random order: the input order svoFromVoxelGrid cannot assume (its radix sort runs)
    perm = torch.randperm(n, device="cuda:%d" % device)
    cen_s, col_s = cen[perm].contiguous(), col[perm].contiguous()
    del perm
    svo_s = pkg.SVO(center, half, D, reserve_nodes=max(1 << 20, int(2.7 * n)), device=device)
    es = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 2)]
    es[0].record()
    for k in range(reps + 1):
        svo_s.integrate_voxels(cen_s, col_s)
        es[k + 1].record()
    svo_s.sync()
    sync()
    ms_s = [es[k].elapsed_time(es[k + 1]) for k in range(reps + 1)]
    svo_s.set_stage_timing(True)
    svo_s.integrate_voxels(cen_s, col_s)
    stage_s = svo_s.stage_times()
    svo_s.set_stage_timing(False)
    cn_s = svo_s.counters()
    same_structure = svo_s.size == nodes
    # full-size property check: the node indices / child pointers do not depend on the input order (word0 of every node;
    # the colours do, by the reference's rule that colour j goes to the j-th smallest key, quirk Q11)
    same_word0 = False
    if same_structure:
        pa, pb = pkg.shard.pool_tensor(svo, nodes), pkg.shard.pool_tensor(svo_s, nodes)
        same_word0 = bool(torch.equal(pa[0::2], pb[0::2]))
        del pa, pb
    svo_s.close()
    del cen_s, col_s
    # ONE map built by ALL ranks from this grid (shard.integrate_voxels_sharded: Morton-range slices, one all-gather of the
    # per-pass split counters, one of the changes): strong scaling of svoFromVoxelGrid, pool identical to the one above
    sharded = None
    if world > 1:
        import torch.distributed as dist
        svo_sh = pkg.SVO(center, half, D, reserve_nodes=max(1 << 20, int(2.7 * n)), device=device)
        times = []
        for k in range(reps + 1):
            sync()
            dist.barrier()
            t0 = time.perf_counter()
            pkg.shard.integrate_voxels_sharded(svo_sh, cen, col)
            sync()
            dist.barrier()
            times.append((time.perf_counter() - t0) * 1e3)
        same = svo_sh.size == nodes
        if same and n <= 80_000_000:  # (compare on the device: both pools after reps + 1 observations)
            pa, na, _, _ = svo.view()
            pb, nb, _, _ = svo_sh.view()
            ta = torch.as_tensor(pkg.shard._DeviceAlias(pa, 2 * na), device="cuda:%d" % device)
            tb = torch.as_tensor(pkg.shard._DeviceAlias(pb, 2 * nb), device="cuda:%d" % device)
            same = bool(torch.equal(ta, tb))
        sharded = {"ranks": world, "first_ms": times[0], "steady_ms": float(np.median(times[1:])),
                   "single_gpu_first_ms": ms[0], "single_gpu_steady_ms": steady_ms,
                   "pool_identical_to_single_gpu_build": bool(same),
                   "timing": "host wall clock between barriers (device synchronised), all ranks"}
        svo_sh.close()
    # the reference's renderer only shows nodes whose occupancy counter has saturated (a sample terminates a ray when
    # alpha >= 254, cone_tracing_kernels.cu:115-121; a once-observed surface is transparent): observe the grid 64 times
    for _ in range(max(0, 64 - (reps + 1))):
        svo.integrate_voxels(cen, col)
    svo.sync()
    # ---- raycast: camera outside the cube for the bunny (2.5 half edges from the centre), inside the
    # atrium for sponza; interleaved row bands over the ranks
    c = np.asarray(center, dtype=np.float64)
    if name == "cfg2":
        view = look_at(c + np.array([0.0, 0.0, 2.5 * half]), c)
    else:
        view = look_at(c + np.array([-0.55 * half, -0.12 * half, 0.0]), c + np.array([0.6 * half, 0.0, 0.05 * half]))
    band = pkg.shard.band_height(H, world)
    rows = sum(r for _, r in pkg.shard.row_bands(H, world, rank, band))
    out = torch.empty((max(rows, 1), W, 4), dtype=torch.uint8, device="cuda:%d" % device)
    for _ in range(2):
        svo.raycast_bands(out, W, H, band, world, rank, 45.0, view)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    r0.record()
    for _ in range(5):
        svo.raycast_bands(out, W, H, band, world, rank, 45.0, view)
    r1.record()
    sync()
    ray_ms = r0.elapsed_time(r1) / 5
    st = pkg.RaycastStats()
    lit = None
    if world == 1:
        img = svo.raycast(W, H, 45.0, view, stats=st)
        lit = int(np.count_nonzero(img[..., :3].any(axis=2)))
    res = {
        "workload": "%s: %s (%d triangles%s) -> depth-%d SVO, %dx%d raycast" % (
            name, asset_name if not stand_in else "STAND-IN icosphere (asset not staged)", T.shape[0],
            ", vertices x %.4g" % scale if scale != 1.0 else "", D, W, H),
        "named_asset": not stand_in, "depth": D, "half_edge": half, "voxels": n, "nodes": int(nodes),
        "pool_GB": nodes * 8 / 1e9,
        "voxelise_ms": {"first_call": vox_first * 1e3, "second_call": vox_ms},
        "integrate_voxels": {
            "morton_ordered_input": {"first_ms": ms[0], "steady_ms": steady_ms,
                                     "algorithmic_bytes": int(cn.algorithmic_bytes),
                                     "GBps": cn.algorithmic_bytes / (steady_ms / 1e3) / 1e9,
                                     "frac_of_hbm_peak": cn.algorithmic_bytes / (steady_ms / 1e3) / 1e9 / peak},
            "shuffled_input": {"first_ms": ms_s[0], "steady_ms": float(np.median(ms_s[1:])),
                               "algorithmic_bytes": int(cn_s.algorithmic_bytes),
                               "GBps": cn_s.algorithmic_bytes / (float(np.median(ms_s[1:])) / 1e3) / 1e9,
                               "frac_of_hbm_peak": cn_s.algorithmic_bytes / (float(np.median(ms_s[1:])) / 1e3) / 1e9 / peak,
                               "stage_ms": {"k_emit": stage_s[0], "k_sort": stage_s[1], "k_structure": stage_s[2],
                                            "k_levels": stage_s[3]},
                               "same_node_count_as_ordered": bool(same_structure),
                               "same_child_pointers_as_ordered": same_word0},
            "U": int(cn.n_unique), "peak_GBps": peak},
        "sharded_build": sharded,
        "raycast": {"res": [W, H], "ranks": world, "ms_this_rank": ray_ms, "rows_this_rank": rows,
                    "mrays_per_s_this_rank": W * rows / (ray_ms / 1e3) / 1e6,
                    "steps_per_ray": (st.steps / float(st.rays)) if st.rays else None, "lit_pixels": lit},
    }
    svo.close()
    return res, ray_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res, ray_ms = run(a.config, a.depth, a.reps, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ray_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ray_ms = float(t.item())
        W, H = res["raycast"]["res"]
        res["raycast"]["ms_max_over_ranks"] = ray_ms
        res["raycast"]["mrays_per_s"] = W * H / (ray_ms / 1e3) / 1e6
        res["raycast"]["tree"] = "every rank voxelises and builds its own replica of the map (no exchange)"
        dist.barrier()
        dist.destroy_process_group()
    else:
        W, H = res["raycast"]["res"]
        res["raycast"]["mrays_per_s"] = W * H / (ray_ms / 1e3) / 1e6
    if rank == 0:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
