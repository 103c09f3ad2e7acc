"""Worker of tests/test_gpu_multirank.py (one process per GPU, launched by torch.distributed.run): replicas of one map
and the row-band raycast, checked bit for bit against rank 0's single-GPU result."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import pkg, view_for_pose  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = pkg()
    S = P.shard
    D, w, h = 12, 320, 240
    center, half = P.synth.tree_params(D)
    fx, fy = P.synth.focal(w, h)
    svo = P.SVO(center, half, D, device=local)
    out = {"rank": rank}
    # 1. rank 0 fuses frames; the others follow by full copy (frame 0) and by deltas (every later frame)
    pose = None
    delta_sizes = []
    for k in range(6):
        pose = P.synth.orbit_pose(10 * k)
        if rank == 0:
            depth, rgb = P.synth.make_frame(w, h, pose, seed=k)
            svo.integrate_depth(depth, rgb, fx, fy, pose)
            svo.sync()
        if k == 0:
            n = S.replicate_tree(svo, 0, device="cuda")
            out["full_copy_nodes"] = n
        else:
            delta_sizes.append(S.replicate_delta(svo, 0))
    out["delta_bytes"] = delta_sizes
    mine = torch.from_numpy(svo.pool().view(np.int32)).cuda()
    n_words = torch.tensor([mine.numel()], dtype=torch.int64, device="cuda")
    sizes = [torch.zeros_like(n_words) for _ in range(world)]
    dist.all_gather(sizes, n_words)
    out["sizes_equal"] = bool(all(int(s.item()) == int(n_words.item()) for s in sizes))
    ref = mine.clone()
    dist.broadcast(ref, 0)
    out["pool_equal_rank0"] = bool(torch.equal(ref, mine)) if out["sizes_equal"] else False
    # 2. a replica integrates on its own after the delta stream (its device-side size / caches are consistent): the same
    #    frame on every rank must give the same pool again
    pose7 = P.synth.orbit_pose(70)
    depth, rgb = P.synth.make_frame(w, h, pose7, seed=7)
    svo.integrate_depth(depth, rgb, fx, fy, pose7)
    mine = torch.from_numpy(svo.pool().view(np.int32)).cuda()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    out["pool_equal_after_local_frame"] = bool(ref.numel() == mine.numel() and torch.equal(ref, mine))
    # 3. raycast in interleaved row bands (default and a non-default band height), gathered on rank 0, against rank 0's
    #    single-GPU image
    view = view_for_pose(pose7)
    ok = True
    for W, H, band in ((640, 480, None), (333, 217, 5), (1920, 1080, 16)):
        bh = band or S.band_height(H, world)
        bands = S.row_bands(H, world, rank, bh)
        rows = sum(r for _, r in bands)
        buf = torch.empty((max(rows, 1), W, 4), dtype=torch.uint8, device="cuda")
        got_rows = svo.raycast_bands(buf, W, H, bh, world, rank, 45.0, view)
        torch.cuda.synchronize()
        assert got_rows == rows
        tiles, off = [], 0
        for _, r in bands:
            tiles.append(buf[off:off + r])
            off += r
        img = S.gather_image(bands, tiles, H, W, dst=0, band=bh)
        if rank == 0:
            want = torch.from_numpy(svo.raycast(W, H, 45.0, view)).cuda()
            ok = ok and bool(torch.equal(img, want))
    out["image_equal"] = ok
    # 4. a replica with another geometry must refuse the pool
    other = P.SVO(center, half * 2.0, D, device=local)
    refused = True
    if rank != 0:
        try:
            other.reserve(svo.size)
            other.adopt(svo.size, D, center, half)
            refused = False
        except P.OslError:
            refused = True
    out["geometry_mismatch_refused"] = refused
    # 5. ONE map built by all ranks from ONE voxel grid (Morton-range shards, split-count prefix over the ranks): pools
    #    bit-identical on every rank with the single-GPU build, also for the second observation (Q3 leaf splits, blends
    #    with non-empty leaves) and for a second, overlapping grid
    Dm = 9
    V, T = P.synth.icosphere(4, 0.8, (0.05, -0.02, 0.1))
    rng = np.random.default_rng(5)
    colors = rng.uniform(0.1, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    cen, col = P.meshToVoxelGrid(V, T, colors, (0.0, 0.0, 0.0), 1.0, Dm, device=local)
    V2, T2 = P.synth.icosphere(3, 0.55, (-0.2, 0.1, 0.0))
    cen2, col2 = P.meshToVoxelGrid(V2, T2, colors[:T2.shape[0]], (0.0, 0.0, 0.0), 1.0, Dm, device=local)
    single = P.SVO((0.0, 0.0, 0.0), 1.0, Dm, device=local)
    sharded = P.SVO((0.0, 0.0, 0.0), 1.0, Dm, device=local)
    ok_shard, sizes_shard = True, []
    for c4, k4 in ((cen, col), (cen, col), (cen2, col2), (cen, col)):
        single.integrate_voxels(c4, k4)
        S.integrate_voxels_sharded(sharded, c4, k4)
        a, b = single.pool(), sharded.pool()
        ok_shard = ok_shard and a.size == b.size and bool(np.array_equal(a, b))
        sizes_shard.append(int(a.size // 2))
    out["sharded_build_equal"] = ok_shard
    out["sharded_build_nodes"] = sizes_shard
    out["sharded_voxels"] = [int(cen.shape[0]), int(cen2.shape[0])]
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("MR_RESULT " + json.dumps(gathered))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
