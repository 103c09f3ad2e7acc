#!/usr/bin/env python
"""Pinned host -> device copy rate for the e2e frame payload (614 400 B depth + 921 600 B colour per frame)."""
import torch


def main():
    n = 64
    hd = [torch.empty(614400, dtype=torch.uint8).pin_memory() for _ in range(n)]
    hc = [torch.empty(921600, dtype=torch.uint8).pin_memory() for _ in range(n)]
    dd = [torch.empty(614400, dtype=torch.uint8, device="cuda") for _ in range(8)]
    dc = [torch.empty(921600, dtype=torch.uint8, device="cuda") for _ in range(8)]
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        for rep in range(2):
            e0.record(s)
            for k in range(2000):
                dd[k % 8].copy_(hd[k % n], non_blocking=True)
                dc[k % 8].copy_(hc[k % n], non_blocking=True)
            e1.record(s)
            s.synchronize()
    us = e0.elapsed_time(e1) / 2000 * 1e3
    print("two copies per frame (1 536 000 B): %.1f us/frame = %.1f GB/s -> PCIe bound of the e2e path: %.0f frames/s"
          % (us, 1.536e6 / us / 1e3, 1e6 / us))
    # the same bytes over TWO streams (depth on one, colour on the other): do two copy engines raise the rate?
    s2 = torch.cuda.Stream()
    f0, f1, g1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for rep in range(2):
        torch.cuda.synchronize()
        f0.record(s)
        s2.wait_event(f0)
        for k in range(2000):
            with torch.cuda.stream(s):
                dd[k % 8].copy_(hd[k % n], non_blocking=True)
            with torch.cuda.stream(s2):
                dc[k % 8].copy_(hc[k % n], non_blocking=True)
        f1.record(s)
        g1.record(s2)
        torch.cuda.synchronize()
    us2 = max(f0.elapsed_time(f1), f0.elapsed_time(g1)) / 2000 * 1e3
    print("same over two streams: %.1f us/frame = %.1f GB/s" % (us2, 1.536e6 / us2 / 1e3))
    # one 8 MB copy: the link's large-transfer rate
    big_h = torch.empty(8 << 20, dtype=torch.uint8).pin_memory()
    big_d = torch.empty(8 << 20, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(s):
        for rep in range(2):
            e0.record(s)
            for k in range(100):
                big_d.copy_(big_h, non_blocking=True)
            e1.record(s)
            s.synchronize()
    print("8 MB copies: %.1f GB/s" % ((8 << 20) * 100 / (e0.elapsed_time(e1) * 1e-3) / 1e9))


if __name__ == "__main__":
    main()
