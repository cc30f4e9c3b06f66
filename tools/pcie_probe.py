#!/usr/bin/env python3
"""PCIe copy bandwidth probe (pinned host memory): D2H, H2D, H2D split over k streams, both directions at once."""
import time

import torch


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    n_dn, n_up = 50_000_000, 20_000_000
    h_dn = torch.empty(n_dn, dtype=torch.uint8).pin_memory()
    h_up = torch.empty(n_up, dtype=torch.uint8).pin_memory()
    d_dn = torch.empty(n_dn, dtype=torch.uint8, device="cuda")
    d_up = torch.empty(n_up, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(4)]
    t = timed(lambda: h_dn.copy_(d_dn, non_blocking=True))
    print("D2H 50 MB, 1 stream : %.3f ms  %.1f GB/s" % (t * 1e3, n_dn / t / 1e9))
    for k in (1, 2, 4):
        def up(k=k):
            c = n_up // k
            for i in range(k):
                with torch.cuda.stream(streams[i]):
                    d_up[i * c:(i + 1) * c].copy_(h_up[i * c:(i + 1) * c], non_blocking=True)
        t = timed(up)
        print("H2D 20 MB, %d streams: %.3f ms  %.1f GB/s" % (k, t * 1e3, n_up / t / 1e9))

    def both():
        with torch.cuda.stream(streams[0]):
            h_dn.copy_(d_dn, non_blocking=True)
        with torch.cuda.stream(streams[1]):
            d_up.copy_(h_up, non_blocking=True)
    t = timed(both)
    print("both directions     : %.3f ms" % (t * 1e3))


if __name__ == "__main__":
    main()
