"""Multi-GPU decomposition of the hot path, one process per GPU over torch.distributed.

  integrate  one map per rank: every rank fuses its own RGB-D stream ("replicas", weak scaling, no data-path
             collective).  Frames of ONE incremental map cannot be sharded over ranks (alpha += 2 and the colour blend
             are order dependent per leaf); the spatially sharded single map needs only the exclusive prefix of the
             per-pass split counts over ranks (`split_count_prefix`, the north-star's "NCCL only for the per-level
             node-count prefix") and is designed in DESIGN.md section 7.
  raycast    rays are independent: image rows are dealt to ranks in interleaved bands (`row_bands`), every rank holds
             the whole tree (`replicate_tree`: broadcast of the flat 2*n uint32 pool, the reference's own wire format,
             octree.cpp:113-169), `gather_image` assembles the picture on one rank.

Works with the NCCL backend (CUDA tensors) and with gloo (CPU tensors; used by the CPU test-suite)."""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def row_bands(h, world, rank, band=None):
    """Rows of an h-row image owned by `rank`: interleaved bands of `band` rows (default: about 4 bands per rank, so
    that sky / near-geometry imbalance averages out).  Returns a list of (row0, rows); over all ranks the bands
    partition [0, h)."""
    if world <= 1:
        return [(0, h)]
    if band is None:
        band = max(1, -(-h // (4 * world)))
    out = []
    k = 0
    for row0 in range(0, h, band):
        if k % world == rank:
            out.append((row0, min(band, h - row0)))
        k += 1
    return out


def band_height(h, world):
    """band height `row_bands` uses by default (about 4 bands per rank)"""
    return h if world <= 1 else max(1, -(-h // (4 * world)))


def split_count_prefix(counts, group=None):
    """counts: int64 tensor [P] of this rank's per-pass split counts (|codes[i]| restricted to the rank's key range;
    rank order = key order).  Returns (base, total): base[i] = number of tiles allocated in pass i by lower ranks plus
    everything allocated in earlier passes by ANY rank -- the rank's first global tile rank in pass i, reproducing the
    reference's allocation order (pass, then numeric key; svo.cu:220,284) -- and total = tiles allocated by all ranks."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(gathered, counts, group=group)
    allc = torch.stack(gathered)                      # [world, P]
    per_pass = allc.sum(dim=0)                        # [P]
    pass_base = torch.cumsum(per_pass, 0) - per_pass  # exclusive over passes
    lower = allc[:rank].sum(dim=0) if rank > 0 else torch.zeros_like(counts)
    return pass_base + lower, int(per_pass.sum().item())


class _DeviceAlias:
    """a raw device pointer as a torch tensor (no copy) through __cuda_array_interface__"""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2}


def pool_tensor(svo, n_nodes):
    """The first n_nodes nodes of `svo`'s pool as an int32 CUDA tensor ALIASING the pool (osl_svo_pool_device): a
    collective can read from / write into the tree without a staging copy."""
    import torch
    ptr, cap = svo.pool_device()
    assert n_nodes <= cap
    return torch.as_tensor(_DeviceAlias(ptr, 2 * n_nodes), device="cuda:%d" % svo.device)


def replicate_tree(svo, src=0, group=None, device=None):
    """Broadcast the node pool of rank `src` to every rank (all ranks then hold the same tree).

    Trees on GPUs (objects with `pool_device`): the collective reads the source's pool and writes straight into the
    receivers' pools -- device memory to device memory over NVLink, no host copy -- after a small header with the
    geometry (max depth, centre, half edge: what a node index means) has been compared on every rank; the receiver
    validates the child pointers before it publishes the tree (osl_svo_adopt).  Anything else (the CPU test doubles of
    the gloo suite) goes through `.pool()` / `.load()` host arrays."""
    import torch
    dist = _dist()
    rank = dist.get_rank(group)
    on_gpu = hasattr(svo, "pool_device") and device is not None and str(device).startswith("cuda")
    if not on_gpu:
        n = torch.zeros(1, dtype=torch.int64, device=device)
        pool = None
        if rank == src:
            pool = np.ascontiguousarray(svo.pool(), dtype=np.uint32)
            n[0] = pool.size
        dist.broadcast(n, src, group=group)
        words = int(n.item())
        buf = torch.empty(words, dtype=torch.int32, device=device)
        if rank == src:
            buf.copy_(torch.from_numpy(pool.view(np.int32)))
        dist.broadcast(buf, src, group=group)
        if rank != src:
            svo.load(buf.cpu().numpy().view(np.uint32))
        return words // 2
    dev = "cuda:%d" % svo.device
    hdr = torch.zeros(6, dtype=torch.float64, device=dev)
    if rank == src:
        _, n_nodes, center, half = svo.view()
        hdr[:] = torch.tensor([n_nodes, svo.max_depth, center[0], center[1], center[2], half], dtype=torch.float64)
    dist.broadcast(hdr, src, group=group)
    h = hdr.cpu().tolist()
    n_nodes, depth, center, half = int(h[0]), int(h[1]), (h[2], h[3], h[4]), h[5]
    if n_nodes == 0:
        if rank != src:
            svo.reset()
        return 0
    if rank != src:
        svo.reserve(n_nodes)
    dist.broadcast(pool_tensor(svo, n_nodes), src, group=group)  # NCCL: pool to pool
    torch.cuda.synchronize()
    if rank != src:
        svo.adopt(n_nodes, depth, center, half)  # raises when the geometry differs or the pool is corrupt
    return n_nodes


def replicate_delta(svo, src=0, group=None):
    """After ONE integrate call on rank `src` (and none on the others): ship what that call changed -- the touched
    nodes' (index, word0, word1) and the appended tiles, osl_svo_delta_pack -- and apply it on every other rank.
    Device to device; ~0.3 MB for a 640x480 frame instead of the whole pool.  Returns the delta's size in bytes."""
    import torch
    dist = _dist()
    rank = dist.get_rank(group)
    dev = "cuda:%d" % svo.device
    nbytes = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        nbytes[0] = svo.delta_bytes()
    dist.broadcast(nbytes, src, group=group)
    n = int(nbytes.item())
    buf = torch.empty((n + 3) // 4, dtype=torch.int32, device=dev)
    if rank == src:
        svo.delta_pack(buf, n)
    dist.broadcast(buf, src, group=group)
    if rank != src:
        torch.cuda.synchronize()
        svo.delta_apply(buf, n)
    return n


def integrate_voxels_sharded(svo, centers4, colors4, group=None):
    """svoFromVoxelGrid of ONE grid into ONE map by all ranks.  `centers4` / `colors4`: the whole grid
    as CUDA tensors [n, 4] on every rank, in Morton order without invalid voxels (the voxelisers' output order); every
    rank's `svo` in the same state.  Rank r analyses the slice [r n / R, (r+1) n / R); ONE all-gather of the per-pass
    split counters (the north-star's "NCCL only for the per-level node-count prefix") turns local tile ranks into the
    reference's global node indices (pass = depth - frontier depth, then numeric key, svo.cu:220,284); a second
    all-gather ships what every rank changed, so that all replicas end up bit-identical with a single-GPU
    osl_integrate_voxels of the whole grid.  Returns (n_counters, delta bytes of this rank)."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = centers4.device
    n = int(centers4.shape[0])
    lo, hi = (rank * n) // world, ((rank + 1) * n) // world
    tot = svo.shard_analyze(centers4, n, lo, hi)
    t_dev = torch.from_numpy(tot.astype(np.int64)).to(dev)
    gathered = [torch.empty_like(t_dev) for _ in range(world)]
    dist.all_gather(gathered, t_dev, group=group)
    allc = torch.stack(gathered).cpu().numpy()                     # [world, NC]
    base = allc[:rank].sum(axis=0).astype(np.uint32) if rank else np.zeros(allc.shape[1], dtype=np.uint32)
    svo.shard_assign(colors4, base, allc.sum(axis=0).astype(np.uint32))
    # what every rank changed: (index, word0, word1) triples of its level lists
    mine = svo.shard_delta_bytes()
    sz = torch.tensor([mine], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(sz) for _ in range(world)]
    dist.all_gather(sizes, sz, group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = (max(sizes) + 3) // 4
    buf = torch.zeros(cap, dtype=torch.int32, device=dev)
    svo.shard_delta_pack(buf, cap * 4)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    torch.cuda.synchronize(dev)
    for q in range(world):
        if q != rank:
            svo.shard_delta_apply(bufs[q].data_ptr(), sizes[q])
    svo.shard_fixup(centers4, n, [(q * n) // world for q in range(1, world)])
    return allc.shape[1], mine


def gather_image(bands, tiles, h, w, dst=0, group=None, band=None):
    """bands: this rank's [(row0, rows)], tiles: matching list of uint8 tensors [rows, w, 4]; `band` = the band height
    EVERY rank used with row_bands (None = row_bands' default).  It is passed explicitly, never inferred from the local
    bands: a rank that owns 0 or 1 bands cannot tell, and mismatched layouts hang the all-gather.  Returns the
    h x w x 4 image on rank `dst` (None elsewhere)."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if band is None:
        band = band_height(h, world)
    assert [tuple(b) for b in bands] == row_bands(h, world, rank, band), "bands do not match the declared band height"
    dev = tiles[0].device if tiles else torch.device("cpu")
    mine = torch.cat([t.reshape(-1) for t in tiles]) if tiles else torch.empty(0, dtype=torch.uint8, device=dev)
    layout = [row_bands(h, world, k, band) for k in range(world)]
    sizes = [sum(r for _, r in layout[k]) * w * 4 for k in range(world)]
    cap = max(sizes)
    padded = torch.zeros(cap, dtype=torch.uint8, device=dev)
    padded[:mine.numel()] = mine
    gathered = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    if rank != dst:
        return None
    img = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
    for k in range(world):
        off = 0
        for row0, rows in layout[k]:
            img[row0:row0 + rows] = gathered[k][off:off + rows * w * 4].reshape(rows, w, 4)
            off += rows * w * 4
    return img
