#!/usr/bin/env python
"""Aggregate pinned D2H bandwidth with one process per GPU, with and without pinning each process
to the CPUs NVML reports as local to its GPU (development tool for the multi-GPU e2e path)."""
import os
import sys
import time

import torch
import torch.multiprocessing as mp


def worker(rank, world, affine, q):
    if False:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception as e:  # noqa: BLE001
            print("affinity failed:", e)
    torch.cuda.set_device(rank)
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.zero_()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    q.put(("ready", rank))
    time.sleep(1.0)
    h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        h.copy_(d, non_blocking=True)
        if affine:                                   # second pass of main(): add concurrent H2D (1/4 of the D2H bytes)
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    q.put(("bw", rank, 4 * n / dt / 1e9, sorted(os.sched_getaffinity(0))[:4]))


def main():
    world = torch.cuda.device_count()
    for affine in (False, True):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        ps = [ctx.Process(target=worker, args=(r, world, affine, q)) for r in range(world)]
        for p in ps:
            p.start()
        res = []
        while len(res) < world:
            m = q.get()
            if m[0] == "bw":
                res.append(m[1:])
        for p in ps:
            p.join()
        res.sort()
        print("d2h+h2d(1/4)" if affine else "d2h only", "total %.1f GB/s" % sum(r[1] for r in res), [round(r[1], 1) for r in res], res[0][2], res[-1][2])


if __name__ == "__main__":
    sys.exit(main())
