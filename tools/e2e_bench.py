"""End-to-end (host buffers -> host buffers) decode and encode throughput of one flag family through the
host-batch C-ABI calls, pinned memory.  usage: e2e_bench.py [--flags 0,1,4] [--blocks 4096] [--reps 3]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import htscodecs_b200 as hb
from htscodecs_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--flags", default="0,1,4")
ap.add_argument("--blocks", type=int, default=4096)
ap.add_argument("--distinct", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--gen", default="qual")
args = ap.parse_args()
n = 1 << 20
ctx = hb.Context(0)
lib = hb.load_library()
blocks = [synth.GENERATORS[args.gen](i, n) for i in range(args.distinct)]
nblk = args.blocks
for f in [int(x, 0) for x in args.flags.split(",")]:
    legacy = bool(f & hb.ORDER_RANS4x8)
    comps, st = ctx.compress_many([b.tobytes() for b in blocks], [f] * args.distinct)
    assert (st == 0).all()
    in_len = np.array([len(comps[i % args.distinct]) for i in range(nblk)], np.uint32)
    in_off = np.zeros(nblk, np.uint64); in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64))
    c_bytes = int(in_len.astype(np.uint64).sum())
    pin_c = hb.PinnedArray(c_bytes + 64)
    for i in range(nblk):
        pin_c.array[int(in_off[i]): int(in_off[i]) + int(in_len[i])] = np.frombuffer(comps[i % args.distinct], np.uint8)
    pin_u = hb.PinnedArray(nblk * n + 64)
    u_off = np.arange(nblk, dtype=np.uint64) * n
    out_len = np.full(nblk, n, np.uint32); status = np.zeros(nblk, np.int32)
    method = np.full(nblk, 1 if legacy else 0, np.uint8)
    td = []
    for r in range(args.reps + 1):
        out_len[:] = n
        t0 = time.perf_counter()
        ctx.uncompress_batch_host(nblk, pin_c.array, in_off, in_len, pin_u.array, u_off, out_len, status, method)
        td.append(time.perf_counter() - t0)
    assert (status == 0).all()
    for i in (0, nblk - 1):
        assert np.array_equal(pin_u.array[i * n:(i + 1) * n], blocks[i % args.distinct])
    # encode: raw (pin_u) -> compressed
    bound = lib.hts_b200_compress_bound_4x8(n) if legacy else hb.rans_compress_bound_4x16(n, f)
    cap = (bound + 15) // 16 * 16
    pin_o = hb.PinnedArray(nblk * cap + 64)
    o_off = np.arange(nblk, dtype=np.uint64) * cap
    raw_len = np.full(nblk, n, np.uint32); order = np.full(nblk, f, np.int32)
    te = []
    for r in range(args.reps + 1):
        o_len = np.full(nblk, cap, np.uint32)
        t0 = time.perf_counter()
        ctx.compress_batch_host(nblk, pin_u.array, u_off, raw_len, pin_o.array, o_off, o_len, status, order)
        te.append(time.perf_counter() - t0)
    assert (status == 0).all() and bytes(pin_o.array[:int(o_len[0])]) == comps[0]
    gb = nblk * n / 1e9
    print(f"flags {f:#x}: e2e decode {gb / min(td[1:]):6.1f} GB/s   e2e encode {gb / min(te[1:]):6.1f} GB/s   (ratio {c_bytes / (nblk * n):.3f})", flush=True)
    del pin_c, pin_u, pin_o
