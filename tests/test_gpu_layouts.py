"""Host-buffer batch calls with blocks scattered through the caller's arenas (gaps, arbitrary order,
unaligned offsets): the staging path that packs blocks instead of mirroring a dense layout.  Bytes
between the blocks must stay untouched."""
import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import synth

pytestmark = pytest.mark.gpu


def _scatter(rng, sizes, gap_lo=70, gap_hi=5000):
    """Offsets for regions of the given sizes, visited in a random order, with random gaps."""
    order = rng.permutation(len(sizes))
    off = np.zeros(len(sizes), np.uint64)
    pos = int(rng.integers(1, 50))
    for i in order:
        off[i] = pos
        pos += int(sizes[i]) + int(rng.integers(gap_lo, gap_hi))
    return off, pos


def test_scattered_blocks_both_directions(oracle):
    rng = np.random.default_rng(3)
    ctx = hb.Context(0)
    gens = ["qual", "tag", "acgt", "wide", "u32", "random"]
    flags = [0, 1, 4, 5, 0x40, 0x81, 0xc5, 9, 0x20]
    raw = [synth.GENERATORS[gens[i % 6]](i, int(rng.integers(1, 40000)) // 4 * 4 + 4).tobytes() for i in range(40)]
    orders = [flags[i % len(flags)] for i in range(40)]
    want = [oracle.compress(d, f) for d, f in zip(raw, orders)]
    n = len(raw)

    # ---- decode: compressed streams scattered in one arena, outputs scattered in another
    in_len = np.array([len(c) for c in want], np.uint32)
    in_off, in_total = _scatter(rng, in_len)
    out_cap = np.array([len(d) for d in raw], np.uint32)
    out_off, out_total = _scatter(rng, out_cap)
    ib = np.full(in_total + 64, 0xAB, np.uint8)
    for i, c in enumerate(want):
        ib[int(in_off[i]): int(in_off[i]) + len(c)] = np.frombuffer(c, np.uint8)
    ob = np.full(out_total + 64, 0xCD, np.uint8)
    out_len = out_cap.copy()
    status = np.zeros(n, np.int32)
    ctx.uncompress_batch_host(n, ib, in_off, in_len, ob, out_off, out_len, status)
    assert (status == 0).all() and (out_len == out_cap).all()
    covered = np.zeros(len(ob), bool)
    for i, d in enumerate(raw):
        a = int(out_off[i])
        assert bytes(ob[a: a + len(d)]) == d, i
        covered[a: a + len(d)] = True
    assert (ob[~covered] == 0xCD).all(), "bytes outside the blocks were written"

    # ---- encode: raw blocks scattered, output regions (capacity = bound) scattered
    r_len = out_cap
    r_off, r_total = _scatter(rng, r_len)
    rb = np.full(r_total + 64, 0x11, np.uint8)
    for i, d in enumerate(raw):
        rb[int(r_off[i]): int(r_off[i]) + len(d)] = np.frombuffer(d, np.uint8)
    caps = np.array([hb.rans_compress_bound_4x16(len(d), f) for d, f in zip(raw, orders)], np.uint32)
    c_off, c_total = _scatter(rng, caps)
    cb = np.full(c_total + 64, 0xEE, np.uint8)
    c_len = caps.copy()
    ctx.compress_batch_host(n, rb, r_off, r_len, cb, c_off, c_len, status, np.array(orders, np.int32))
    assert (status == 0).all()
    inside = np.zeros(len(cb), bool)
    for i, c in enumerate(want):
        a = int(c_off[i])
        assert int(c_len[i]) == len(c) and bytes(cb[a: a + len(c)]) == c, (i, hex(orders[i]))
        inside[a: a + int(caps[i])] = True
    assert (cb[~inside] == 0xEE).all(), "bytes outside the output regions were written"
    ctx.close()
