"""One end-to-end decode of 4096 x 1 MiB blocks with HTSCODECS_B200_TRACE=1: prints the chunk timeline.
usage: HTSCODECS_B200_TRACE=1 python tools/trace_e2e.py [flags]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import htscodecs_b200 as hb
from htscodecs_b200 import synth

flags = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0
nblk, n, distinct = 4096, 1 << 20, 32
raw = [synth.qual_block(i, n).tobytes() for i in range(distinct)]
ctx = hb.Context(0)
comps, st = ctx.compress_many(raw, [flags] * distinct)
assert all(c is not None for c in comps)
sizes = np.array([len(c) for c in comps], np.uint32)
stride = int((sizes.max() + 255) // 256 * 256)
h_in = hb.PinnedArray(nblk * stride)
for i in range(nblk):
    c = comps[i % distinct]
    h_in.array[i * stride: i * stride + len(c)] = np.frombuffer(c, np.uint8)
h_out = hb.PinnedArray(nblk * n)
in_off = (np.arange(nblk, dtype=np.uint64) * np.uint64(stride))
in_len = np.array([sizes[i % distinct] for i in range(nblk)], np.uint32)
out_off = (np.arange(nblk, dtype=np.uint64) * np.uint64(n))
status = np.zeros(nblk, np.int32)
if len(sys.argv) > 2 and sys.argv[2] == "enc":
    # the encode direction: 4096 raw blocks in, streams out
    cap = (hb.rans_compress_bound_4x16(n, flags) + 255) // 256 * 256
    h_raw = hb.PinnedArray(nblk * n)
    for i in range(nblk):
        h_raw.array[i * n: (i + 1) * n] = np.frombuffer(raw[i % distinct], np.uint8)
    h_comp = hb.PinnedArray(nblk * cap)
    c_off = (np.arange(nblk, dtype=np.uint64) * np.uint64(cap))
    r_len = np.full(nblk, n, np.uint32)
    order = np.full(nblk, flags, np.int32)
    for rep in range(3):
        c_len = np.full(nblk, cap, np.uint32)
        sys.stderr.write("---- enc rep %d\n" % rep)
        t0 = time.perf_counter()
        ctx.compress_batch_host(nblk, h_raw.array, out_off, r_len, h_comp.array, c_off, c_len, status, order)
        dt = time.perf_counter() - t0
        assert (status == 0).all()
        print("rep", rep, "e2e encode %.1f GB/s (%.1f ms)" % (nblk * n / dt / 1e9, dt * 1e3))
    sys.exit(0)
for rep in range(3):
    out_len = np.full(nblk, n, np.uint32)
    sys.stderr.write("---- rep %d\n" % rep)
    t0 = time.perf_counter()
    ctx.uncompress_batch_host(nblk, h_in.array, in_off, in_len, h_out.array, out_off, out_len, status)
    dt = time.perf_counter() - t0
    assert (status == 0).all()
    print("rep", rep, "e2e decode %.1f GB/s (%.1f ms)" % (nblk * n / dt / 1e9, dt * 1e3))
