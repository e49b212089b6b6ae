#!/usr/bin/env python
"""Host-link ceilings and the multi-device host-buffer call on every GPU of the box, from ONE process
(development tool; its output is summarised under profiles/).

  1. raw pinned copies, one thread + stream per device: D2H alone, H2D alone, both at once (the decode mix:
     H2D bytes = 0.29 x D2H bytes)
  2. hts_b200_uncompress_batch_host_multi on the bench workload (X_32 order-0, 1 MiB blocks), full-duplex per
     device vs phased across devices, with the per-device stage breakdown

usage: multi_probe.py [--blocks-per-gpu 4096] [--reps 3] [--json out.json] [--flags 4]
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def raw_copies(torch, ndev, mib=1024):
    n = mib << 20
    hs, ds, h2, d2, st = [], [], [], [], []
    for d in range(ndev):
        torch.cuda.set_device(d)
        hs.append(torch.empty(n, dtype=torch.uint8).pin_memory())
        ds.append(torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"))
        h2.append(torch.empty(n * 29 // 100, dtype=torch.uint8).pin_memory())
        d2.append(torch.empty(n * 29 // 100, dtype=torch.uint8, device=f"cuda:{d}"))
        st.append((torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))
    res = {}

    def run(mode):
        bar = threading.Barrier(ndev + 1)
        done = [0.0] * ndev

        def work(d):
            torch.cuda.set_device(d)
            bar.wait()
            if mode in ("d2h", "both"):
                with torch.cuda.stream(st[d][0]):
                    hs[d].copy_(ds[d], non_blocking=True)
            if mode in ("h2d", "both"):
                with torch.cuda.stream(st[d][1]):
                    d2[d].copy_(h2[d], non_blocking=True)
            if mode == "h2d_full":
                with torch.cuda.stream(st[d][1]):
                    ds[d].copy_(hs[d], non_blocking=True)
            torch.cuda.synchronize(d)
            done[d] = time.perf_counter()

        th = [threading.Thread(target=work, args=(d,)) for d in range(ndev)]
        for t in th:
            t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        return max(done) - t0

    for mode in ("d2h", "h2d_full", "both"):
        run(mode)
        dt = min(run(mode) for _ in range(2))
        moved = {"d2h": n, "h2d_full": n, "both": n + n * 29 // 100}[mode] * ndev
        res[mode + "_GBs"] = round(moved / dt / 1e9, 1)
    return res


def main():
    import torch
    import htscodecs_b200 as hb
    from htscodecs_b200 import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks-per-gpu", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--flags", default="4")
    ap.add_argument("--json", default="")
    ap.add_argument("--skip-raw", action="store_true")
    args = ap.parse_args()
    ndev = torch.cuda.device_count()
    out = {"ndev": ndev, "cpus": os.cpu_count()}
    if not args.skip_raw:
        out["raw"] = raw_copies(torch, ndev)
        print("raw copies:", out["raw"], flush=True)
    n = 1 << 20
    distinct = 32
    blocks = [synth.qual_block(i, n) for i in range(distinct)]
    ctx = hb.Context(0)
    for f in [int(x, 0) for x in args.flags.split(",")]:
        comps, st = ctx.compress_many([b.tobytes() for b in blocks], [f] * distinct)
        assert (st == 0).all()
        for nd in sorted({1, ndev}):
            devs = list(range(nd))
            nblk = args.blocks_per_gpu * nd
            in_len = np.array([len(comps[i % distinct]) for i in range(nblk)], np.uint32)
            in_off = np.zeros(nblk, np.uint64); in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64))
            c_bytes = int(in_len.astype(np.uint64).sum())
            pin_c = hb.PinnedArray(c_bytes + 64)
            for i in range(nblk):
                pin_c.array[int(in_off[i]): int(in_off[i]) + int(in_len[i])] = np.frombuffer(comps[i % distinct], np.uint8)
            pin_u = hb.PinnedArray(nblk * n + 64)
            u_off = np.arange(nblk, dtype=np.uint64) * n
            status = np.zeros(nblk, np.int32)
            for phased in ((0, 1) if nd > 1 else (0,)):
                hb.multi_set_phased(None if phased is None else bool(phased))
                ts = []
                for r in range(args.reps + 1):
                    out_len = np.full(nblk, n, np.uint32)
                    t0 = time.perf_counter()
                    hb.uncompress_batch_host_multi(devs, nblk, pin_c.array, in_off, in_len, pin_u.array, u_off, out_len, status)
                    ts.append(time.perf_counter() - t0)
                assert (status == 0).all()
                for i in (0, nblk // 2, nblk - 1):
                    assert np.array_equal(pin_u.array[i * n:(i + 1) * n], blocks[i % distinct])
                gbs = nblk * n / min(ts[1:]) / 1e9
                key = f"flags{f:#x}_ndev{nd}_{'auto' if phased is None else 'phased' if phased else 'duplex'}"
                out[key] = {"e2e_decode_GBs": round(gbs, 1), "stats": hb.multi_last_stats()}
                print(key, round(gbs, 1), "GB/s", json.dumps(out[key]["stats"]), flush=True)
            del pin_c, pin_u
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
