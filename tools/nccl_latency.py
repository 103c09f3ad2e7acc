#!/usr/bin/env python
"""Latency of the collective the single-map sharding would need per frame (all-gather of the per-pass split counts,
D*(D+1) int32 per rank) -- to compare with the frame period.  torchrun --nproc-per-node N tools/nccl_latency.py"""
import os

import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    n = 16 * 17
    mine = torch.full((n,), rank, dtype=torch.int32, device="cuda")
    out = torch.empty((world * n,), dtype=torch.int32, device="cuda")
    for _ in range(50):
        dist.all_gather_into_tensor(out, mine)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 2000
    dist.barrier()
    e0.record()
    for _ in range(K):
        dist.all_gather_into_tensor(out, mine)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / K * 1e3
    t = torch.tensor([us], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("all_gather of %d B per rank over %d ranks: %.1f us per call (back to back, device time, max over ranks)"
              % (n * 4, world, t.item()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
