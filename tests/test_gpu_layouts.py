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


@pytest.mark.parametrize("gap", [1, 16, 64])
def test_small_gaps_are_not_written(gap, oracle):
    """Regions separated by a few bytes (alignment padding, per-block headers the caller owns): the staging path may
    read across them but a block only ever writes out_base[out_off[i] .. + capacity)."""
    rng = np.random.default_rng(gap)
    ctx = hb.Context(0)
    raw = [synth.GENERATORS[("qual", "tag", "acgt")[i % 3]](i, 3000 + 8 * i).tobytes() for i in range(24)]
    orders = [(0, 1, 4, 0x40)[i % 4] for i in range(24)]
    want = [oracle.compress(d, f) for d, f in zip(raw, orders)]
    n = len(raw)
    for regular in (False, True):
        # ---- decode
        in_len = np.array([len(c) for c in want], np.uint32)
        out_cap = np.array([len(d) for d in raw], np.uint32)
        if regular:                                           # equal capacities, evenly spaced: the pitched copy
            out_cap[:] = out_cap.max()
        in_off = np.zeros(n, np.uint64); out_off = np.zeros(n, np.uint64)
        in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64) + gap)
        out_off[1:] = np.cumsum(out_cap[:-1].astype(np.uint64) + gap)
        ib = np.full(int(in_off[-1]) + int(in_len[-1]) + gap, 0xAB, np.uint8)
        for i, c in enumerate(want):
            ib[int(in_off[i]): int(in_off[i]) + len(c)] = np.frombuffer(c, np.uint8)
        ob = np.full(int(out_off[-1]) + int(out_cap[-1]) + gap, 0xCD, np.uint8)
        out_len = out_cap.copy()
        status = np.zeros(n, np.int32)
        ctx.uncompress_batch_host(n, ib, in_off, in_len, ob, out_off, out_len, status)
        assert (status == 0).all()
        inside = np.zeros(len(ob), bool)
        for i, d in enumerate(raw):
            a = int(out_off[i])
            assert int(out_len[i]) == len(d) and bytes(ob[a: a + len(d)]) == d, (regular, i)
            inside[a: a + int(out_cap[i])] = True
        assert (ob[~inside] == 0xCD).all(), "decode wrote between the output regions"
        # ---- encode (capacity = bound, or one common capacity so that the rows are regular)
        caps = np.array([hb.rans_compress_bound_4x16(len(d), f) for d, f in zip(raw, orders)], np.uint32)
        if regular:
            caps[:] = caps.max()
        r_len = np.array([len(d) for d in raw], np.uint32)
        r_off = np.zeros(n, np.uint64); c_off = np.zeros(n, np.uint64)
        r_off[1:] = np.cumsum(r_len[:-1].astype(np.uint64) + gap)
        c_off[1:] = np.cumsum(caps[:-1].astype(np.uint64) + gap)
        rb = np.full(int(r_off[-1]) + int(r_len[-1]) + gap, 0x11, np.uint8)
        for i, d in enumerate(raw):
            rb[int(r_off[i]): int(r_off[i]) + len(d)] = np.frombuffer(d, np.uint8)
        for rep in range(2):                                  # the second call uses the predicted-width pitched copy
            cb = np.full(int(c_off[-1]) + int(caps[-1]) + gap, 0xEE, np.uint8)
            c_len = caps.copy()
            ctx.compress_batch_host(n, rb, r_off, r_len, cb, c_off, c_len, status, np.array(orders, np.int32))
            assert (status == 0).all()
            inside = np.zeros(len(cb), bool)
            for i, c in enumerate(want):
                a = int(c_off[i])
                assert int(c_len[i]) == len(c) and bytes(cb[a: a + len(c)]) == c, (regular, rep, i)
                inside[a: a + int(caps[i])] = True
            assert (cb[~inside] == 0xEE).all(), "encode wrote between the output regions"
    ctx.close()


def test_descending_output_regions(oracle):
    """Evenly spaced but DESCENDING regions must not be mistaken for a pitched layout."""
    ctx = hb.Context(0)
    n = 12
    raw = [synth.qual_block(i, 5000).tobytes() for i in range(n)]
    want = [oracle.compress(d, 1) for d in raw]
    cap = hb.rans_compress_bound_4x16(5000, 1)
    r_off = np.arange(n, dtype=np.uint64) * 5000
    c_off = (np.arange(n, dtype=np.uint64)[::-1] * (cap + 32)).copy()
    rb = np.frombuffer(b"".join(raw), np.uint8).copy()
    status = np.zeros(n, np.int32)
    for rep in range(2):
        cb = np.full(n * (cap + 32), 0xEE, np.uint8)
        c_len = np.full(n, cap, np.uint32)
        ctx.compress_batch_host(n, rb, r_off, np.full(n, 5000, np.uint32), cb, c_off, c_len, status, np.full(n, 1, np.int32))
        assert (status == 0).all()
        for i, c in enumerate(want):
            assert bytes(cb[int(c_off[i]): int(c_off[i]) + int(c_len[i])]) == c
    ctx.close()


def test_encode_capacity_below_the_bound_is_refused(oracle):
    """The reference's coders return NULL when *out_size < rans_compress_bound_4x16 (rANS_static4x16pr.c:396-397,
    :706-707); the batched encoders report HTS_B200_ERR_SIZE for that block, write nothing to it and code the rest."""
    import torch
    ctx = hb.Context(0)
    n = 9
    raw = [synth.random_block(i, 6000).tobytes() for i in range(n)]        # incompressible: the stream is > 6000 bytes
    orders = [0, 1, 4, 5, 0x40, 0x80, 9, hb.ORDER_RANS4x8, hb.ORDER_RANS4x8 | 1]
    lib = hb.load_library()
    bound = np.array([lib.hts_b200_compress_bound_4x8(6000) if f & hb.ORDER_RANS4x8 else hb.rans_compress_bound_4x16(6000, f)
                      for f in orders], np.uint32)
    small = [1, 4, 6, 8]
    caps = bound.copy()
    caps[small] = 6001                                                       # would hold the input, not the bound
    c_off = np.zeros(n, np.uint64)
    c_off[1:] = np.cumsum(bound[:-1].astype(np.uint64) + 64)
    r_off = np.arange(n, dtype=np.uint64) * 6000
    rb = np.frombuffer(b"".join(raw), np.uint8).copy()
    r_len = np.full(n, 6000, np.uint32)
    # host-buffer call
    cb = np.full(int(c_off[-1]) + int(bound[-1]) + 64, 0xEE, np.uint8)
    c_len = caps.copy()
    status = np.zeros(n, np.int32)
    ctx.compress_batch_host(n, rb, r_off, r_len, cb, c_off, c_len, status, np.array(orders, np.int32))
    # device-resident call
    d_cb = torch.full((len(cb),), 0xEE, dtype=torch.uint8, device="cuda")
    d_len = torch.from_numpy(caps.view(np.int32).copy()).cuda()
    d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.compress_batch_dev(n, torch.from_numpy(rb).cuda(), torch.from_numpy(r_off.view(np.int64)).cuda(),
                           torch.from_numpy(r_len.view(np.int32)).cuda(), d_cb, torch.from_numpy(c_off.view(np.int64)).cuda(),
                           d_len, d_st, torch.tensor(orders, dtype=torch.int32, device="cuda"))
    for host, got_b, got_len, got_st in ((True, cb, c_len, status),
                                         (False, d_cb.cpu().numpy(), d_len.cpu().numpy().view(np.uint32), d_st.cpu().numpy())):
        inside = np.zeros(len(cb), bool)
        for i, f in enumerate(orders):
            a = int(c_off[i])
            if i in small:
                assert got_st[i] == -2, (host, i, got_st[i])
                # the device writes nothing; the host-buffer call may copy staging bytes, but only into the block's own region
                if host:
                    inside[a: a + int(caps[i])] = True
                continue
            want = oracle.compress_4x8(raw[i], f & 1) if f & hb.ORDER_RANS4x8 else oracle.compress(raw[i], f)
            assert got_st[i] == 0 and bytes(got_b[a: a + int(got_len[i])]) == want, (host, i)
            inside[a: a + int(bound[i])] = True
        assert (got_b[~inside] == 0xEE).all(), ("a refused block (or a neighbour) wrote outside its region", host)
    ctx.close()
