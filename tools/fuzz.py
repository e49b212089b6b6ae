#!/usr/bin/env python
"""Malformed-stream fuzzing of the decoders (development tool; run on a GPU box under `timeout`).

Valid streams of every codec / flag family are damaged (bit flips weighted towards the header and
table bytes, truncation, byte insertion, random tails) and decoded in batches large enough to reach
the high-occupancy kernel variants as well.  A CUDA fault poisons the context, so after every batch a
known-good stream must still decode; statuses and outputs of damaged streams are unspecified."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def damage(rng, c):
    c = bytearray(c)
    kind = int(rng.integers(0, 6))
    hdr = min(len(c), 96)
    if kind == 0:                                   # flips in the header / tables
        for _ in range(int(rng.integers(1, 5))):
            c[int(rng.integers(0, hdr))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1:                                 # flips anywhere
        for _ in range(int(rng.integers(1, 4))):
            c[int(rng.integers(0, len(c)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 2:                                 # truncate
        c = c[: int(rng.integers(0, len(c)))]
    elif kind == 3:                                 # overwrite a header byte with an extreme value
        c[int(rng.integers(0, hdr))] = int(rng.choice([0, 1, 0x7f, 0x80, 0xff]))
    elif kind == 4:                                 # insert a byte
        c.insert(int(rng.integers(0, hdr)), int(rng.integers(0, 256)))
    else:                                           # random tail
        k = int(rng.integers(0, len(c)))
        c[k:] = bytes(rng.integers(0, 256, len(c) - k, dtype=np.uint8))
    return bytes(c)


def main():
    import htscodecs_b200 as hb
    from htscodecs_b200 import synth
    from oracle_lib import Oracle, RefLib
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--batch", type=int, default=6000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    o, ctx = Oracle(), hb.Context(0)
    ref = RefLib() if RefLib.available() else None
    flags = [0, 1, 4, 5, 0x40, 0x41, 0x80, 0x81, 0xc0, 0xc1, 0xc5, 8, 9, 0x0c, 0x0d, 0xc9, 0x208, 0x809, 0x20, 0x10, 0x11]
    base = []
    for i in range(120):
        gen = ["qual", "wide", "tag", "acgt", "u32", "random"][i % 6]
        n = int(rng.integers(1, 6000))
        d = synth.GENERATORS[gen](i, n).tobytes()
        f = flags[i % len(flags)]
        base.append((o.compress(d, f), n, 0))
        base.append((o.compress_4x8(d, i & 1), n, 1))
        if i % 4 == 0:                              # all-256-symbol data: compressed order-1 tables
            d2 = bytes(rng.permutation(np.arange(256).repeat(8)).astype(np.uint8))
            base.append((o.compress(d2, 1 | (i & 4)), len(d2), 0))
    good = synth.qual_block(2, 10000).tobytes()
    good_c = o.compress(good, 5)
    t0, it, nbad = time.time(), 0, 0
    while time.time() - t0 < args.seconds:
        streams, sizes, methods = [], [], []
        for _ in range(args.batch):
            c, n, m = base[int(rng.integers(0, len(base)))]
            streams.append(damage(rng, c) if rng.random() < 0.9 else c)
            sizes.append(n if rng.random() < 0.8 else int(rng.integers(0, 2 * n + 2)))
            methods.append(m)
        out, status = ctx.uncompress_many_dev(streams, sizes, methods) if it % 2 else ctx.uncompress_many(streams, sizes, methods)
        nbad += int((np.asarray(status) != 0).sum())
        chk, st = ctx.uncompress_many([good_c], [len(good)])
        assert st[0] == 0 and chk[0] == good, "context damaged after batch %d" % it
        it += 1
    print("fuzz ok: %d batches of %d streams, %d rejected, %.0f s" % (it, args.batch, nbad, time.time() - t0))


if __name__ == "__main__":
    main()
