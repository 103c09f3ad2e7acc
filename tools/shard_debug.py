#!/usr/bin/env python
"""Diagnostic for the sharded build: after every step, where the sharded pool differs from the single-GPU pool.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_debug.py [depth] [subdiv]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def depth_of_nodes(pool):
    w0 = pool[0::2]
    n = w0.size
    dep = np.zeros(n, dtype=np.int32)
    dep[:8] = 1
    frontier = np.arange(8)
    d = 1
    while frontier.size:
        has = (w0[frontier] & 0x40000000) != 0
        tiles = (w0[frontier[has]] & 0x3FFFFFFF).astype(np.int64)
        tiles = tiles[(tiles >= 8) & (tiles + 8 <= n)]
        nxt = (tiles[:, None] + np.arange(8)[None, :]).ravel()
        d += 1
        dep[nxt] = d
        frontier = nxt
    return dep


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = graft.load_package()
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    sub = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    V, T = P.synth.icosphere(sub, 0.8, (0.05, -0.02, 0.1))
    colors = np.random.default_rng(5).uniform(0.1, 1.0, size=(T.shape[0], 4)).astype(np.float32)
    cen, col = P.meshToVoxelGrid(V, T, colors, (0.0, 0.0, 0.0), 1.0, D, device=local)
    single = P.SVO((0.0, 0.0, 0.0), 1.0, D, device=local)
    sharded = P.SVO((0.0, 0.0, 0.0), 1.0, D, device=local)
    for step in range(3):
        single.integrate_voxels(cen, col)
        P.shard.integrate_voxels_sharded(sharded, cen, col)
        a, b = single.pool(), sharded.pool()
        for r in range(world):
            dist.barrier()
            if r != rank:
                continue
            print("rank %d step %d: voxels %d, nodes single %d sharded %d" % (rank, step, cen.shape[0], a.size // 2, b.size // 2))
            if a.size != b.size:
                continue
            dep = depth_of_nodes(a)
            bad0 = np.flatnonzero(a[0::2] != b[0::2])
            bad1 = np.flatnonzero(a[1::2] != b[1::2])
            print("   word0 differs at %d nodes, word1 at %d nodes" % (bad0.size, bad1.size))
            for name, bad, off in (("word0", bad0, 0), ("word1", bad1, 1)):
                if bad.size:
                    print("   %s by depth:" % name, dict(zip(*np.unique(dep[bad], return_counts=True))))
                    for i in bad[:6]:
                        print("      node %d depth %d single %08x sharded %08x" % (i, dep[i], a[2 * i + off], b[2 * i + off]))
            sys.stdout.flush()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
