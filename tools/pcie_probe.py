#!/usr/bin/env python
"""Measures what the host link of this box can do, so bench.py's e2e figure can be read against
it: pinned H2D, pinned D2H, and both at once, for a few transfer sizes (development tool)."""
import json
import sys
import time

import torch


def main():
    dev = torch.device("cuda", 0)
    res = {}
    for mb in (16, 96, 512, 2048):
        n = mb << 20
        h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
        d_a = torch.empty(n, dtype=torch.uint8, device=dev)
        d_b = torch.empty(n, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        reps = max(2, 4096 // mb // 4)

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        def h2d():
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)

        def both():
            h2d()
            d2h()

        t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
        res[f"{mb}MiB"] = {"h2d_GBs": n / t_h2d / 1e9, "d2h_GBs": n / t_d2h / 1e9,
                           "bidir_each_GBs": n / t_both / 1e9}
        del h_a, h_b, d_a, d_b
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    sys.exit(main())
