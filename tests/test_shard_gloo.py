"""CPU, world_size 2 over gloo: the host logic of the multi-GPU decomposition (octree-slam_b200/shard.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeTree:
    """stands in for SVO on a CPU box: pool() / load()"""

    def __init__(self, pool=None):
        self._pool = pool

    def pool(self):
        return self._pool

    def load(self, pool):
        self._pool = np.array(pool, dtype=np.uint32)


def _worker(rank, world, port, h, w, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = pkg()
        S = P.shard
        # 1. row bands partition the image and every rank renders only its rows
        bands = S.row_bands(h, world, rank)
        ref = np.arange(h * w * 4, dtype=np.uint32).astype(np.uint8).reshape(h, w, 4)  # stands in for the render
        tiles = [torch.from_numpy(ref[r0:r0 + n].copy()) for r0, n in bands]
        img = S.gather_image(bands, tiles, h, w, dst=0)
        ok_img = True if rank != 0 else bool(np.array_equal(img.numpy(), ref))
        # a non-default band height, chosen so that one rank owns a single band (or none): the layout must come from
        # the declared height, not from what a rank can infer from its own bands
        for bh in (h - 1, h // 2 + 1, 5):
            b2 = S.row_bands(h, world, rank, bh)
            t2 = [torch.from_numpy(ref[r0:r0 + n].copy()) for r0, n in b2]
            i2 = S.gather_image(b2, t2, h, w, dst=0, band=bh)
            ok_img = ok_img and (rank != 0 or bool(np.array_equal(i2.numpy(), ref)))
        # 2. pool replication: rank 0's tree reaches rank 1 bit for bit
        rng = np.random.default_rng(3)
        pool0 = rng.integers(0, 2 ** 32, size=2 * 1000, dtype=np.uint64).astype(np.uint32)
        tree = _FakeTree(pool0 if rank == 0 else None)
        n = S.replicate_tree(tree, src=0)
        ok_pool = n == 1000 and np.array_equal(tree.pool(), pool0)
        # 3. per-pass split-count prefix == the serial allocation order (pass, then rank = key range)
        counts = torch.tensor([[5, 0, 7, 1], [2, 3, 0, 4]][rank], dtype=torch.int64)
        base, total = S.split_count_prefix(counts)
        want = [[0, 7, 10, 17], [5, 7, 17, 18]][rank]
        ok_prefix = base.tolist() == want and total == 22
        q.put((rank, ok_img, ok_pool, ok_prefix, [list(b) for b in bands]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("h,w", [(48, 8), (37, 5)])
def test_world_size_2_gloo(h, w):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = []
    for rank, ok_img, ok_pool, ok_prefix, bands in res:
        assert ok_img and ok_pool and ok_prefix, (rank, ok_img, ok_pool, ok_prefix)
        rows += [r for r0, n in bands for r in range(r0, r0 + n)]
    assert sorted(rows) == list(range(h))  # the bands of all ranks partition the image


def test_row_bands_single_rank_and_many_ranks():
    S = pkg().shard
    assert S.row_bands(480, 1, 0) == [(0, 480)]
    for world in (2, 3, 4, 8):
        rows = []
        for r in range(world):
            rows += [y for r0, n in S.row_bands(1080, world, r) for y in range(r0, r0 + n)]
        assert sorted(rows) == list(range(1080))
