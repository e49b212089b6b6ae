#!/usr/bin/env python
"""Device-resident throughput of every codec path (not the driver's bench: a development tool).

    python tools/pathbench.py [--blocks 1024] [--gen qual] [--flags 0,1,4,5,...]
Prints one line per (flags, direction): uncompressed GB/s with CUDA events on the context stream.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import htscodecs_b200 as hb
    from htscodecs_b200 import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1024)
    ap.add_argument("--distinct", type=int, default=32)
    ap.add_argument("--gen", default="qual")
    ap.add_argument("--flags", default="4,5,0,1,0x44,0x85,0xc5,0x0d,0x41,0x81,9")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--size", type=int, default=1 << 20)
    args = ap.parse_args()
    n, nblk = args.size, args.blocks
    ctx = hb.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    blocks = [synth.GENERATORS[args.gen](i, n) for i in range(args.distinct)]
    d_raw = torch.from_numpy(np.concatenate([blocks[i % args.distinct] for i in range(nblk)])).cuda()
    raw_off = torch.arange(nblk, dtype=torch.int64, device="cuda") * n
    raw_len = torch.full((nblk,), n, dtype=torch.int32, device="cuda")
    for fs in args.flags.split(","):
        f = int(fs, 0)
        legacy = bool(f & hb.ORDER_RANS4x8)          # e.g. --flags 0x40000000,0x40000001
        cap = hb.load_library().hts_b200_compress_bound_4x8(n) if legacy else hb.rans_compress_bound_4x16(n, f)
        cap = (cap + 15) // 16 * 16
        d_comp = torch.empty(nblk * cap, dtype=torch.uint8, device="cuda")
        comp_off = torch.arange(nblk, dtype=torch.int64, device="cuda") * cap
        comp_len = torch.full((nblk,), cap, dtype=torch.int32, device="cuda")
        status = torch.zeros(nblk, dtype=torch.int32, device="cuda")
        order = torch.full((nblk,), f, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()

        def enc():
            comp_len.fill_(cap)
            torch.cuda.synchronize()
            ctx.compress_batch_dev(nblk, d_raw, raw_off, raw_len, d_comp, comp_off, comp_len, status, order, sync=False)
        ctx.compress_batch_dev(nblk, d_raw, raw_off, raw_len, d_comp, comp_off, comp_len, status, order, sync=True)   # sizes the arena
        enc(); torch.cuda.synchronize()
        assert int((status != 0).sum()) == 0, "encode failed"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tt = 0.0
        for _ in range(args.reps):
            comp_len.fill_(cap); torch.cuda.synchronize()
            e0.record(stream)
            ctx.compress_batch_dev(nblk, d_raw, raw_off, raw_len, d_comp, comp_off, comp_len, status, order, sync=False)
            e1.record(stream); torch.cuda.synchronize()
            tt += e0.elapsed_time(e1)
        enc_gbs = nblk * n / (tt / args.reps * 1e-3) / 1e9
        csz = int(comp_len.to(torch.int64).sum())
        # decode what we just encoded
        d_out = torch.empty(nblk * n, dtype=torch.uint8, device="cuda")
        out_len = torch.full((nblk,), n, dtype=torch.int32, device="cuda")
        in_len = comp_len.clone()
        method = torch.full((nblk,), 1, dtype=torch.uint8, device="cuda") if legacy else None
        ctx.uncompress_batch_dev(nblk, d_comp, comp_off, in_len, d_out, raw_off, out_len, status, method)
        assert int((status != 0).sum()) == 0, "decode failed"
        assert torch.equal(d_out, d_raw), "round trip mismatch"
        tt = 0.0
        for _ in range(args.reps):
            out_len.fill_(n); torch.cuda.synchronize()
            e0.record(stream)
            ctx.uncompress_batch_dev(nblk, d_comp, comp_off, in_len, d_out, raw_off, out_len, status, method, sync=False)
            e1.record(stream); torch.cuda.synchronize()
            tt += e0.elapsed_time(e1)
        dec_gbs = nblk * n / (tt / args.reps * 1e-3) / 1e9
        print(f"gen={args.gen} flags={f:#010x} ratio={csz / (nblk * n):.3f}  encode {enc_gbs:8.1f} GB/s   decode {dec_gbs:8.1f} GB/s", flush=True)
        del d_comp, d_out


if __name__ == "__main__":
    main()
