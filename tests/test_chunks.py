"""CPU-only: chunk planning of the host-buffer calls (hts_b200_plan_chunks -- pure host logic, no device work).
Chunks must tile the block list in order, stay near the size target, start small so the device->host stream
starts early, and grow with the slowest stream's serial time (a 1 MiB 4-way stream is 8 x the steps of an X_32
one, so its chunks carry 8 x the bytes)."""
import ctypes as C

import numpy as np

import htscodecs_b200 as hb

MiB = 1 << 20


def _plan(enc, in_len, out_len, first_byte=0, method=None, order=None):
    lib = hb.load_library()
    n = len(in_len)
    in_len = np.asarray(in_len, np.uint32)
    out_len = np.asarray(out_len, np.uint32)
    in_off = np.zeros(n, np.uint64)
    in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64))
    base = np.zeros(int(in_len.astype(np.uint64).sum()) + 1, np.uint8)
    fb = np.broadcast_to(np.asarray(first_byte, np.uint8), (n,))
    base[in_off.astype(np.int64)[in_len > 0]] = fb[in_len > 0]
    cuts = np.zeros(n + 2, np.int32)
    m = None if method is None else np.asarray(method, np.uint8)
    o = None if order is None else np.asarray(order, np.int32)
    k = lib.hts_b200_plan_chunks(int(enc), n, base.ctypes.data, in_off.ctypes.data, in_len.ctypes.data,
                                 out_len.ctypes.data, None if m is None else m.ctypes.data,
                                 None if o is None else o.ctypes.data, cuts.ctypes.data, len(cuts))
    assert 2 <= k <= n + 1
    c = cuts[:k]
    assert c[0] == 0 and c[-1] == n and (np.diff(c) > 0).all()
    return c


def _chunk_bytes(c, in_len, out_len):
    w = np.asarray(in_len, np.uint64) + np.asarray(out_len, np.uint64)
    return np.array([int(w[a:b].sum()) for a, b in zip(c[:-1], c[1:])])


def test_small_batch_is_one_chunk():
    c = _plan(False, [30000] * 10, [100000] * 10, first_byte=4)
    assert list(c) == [0, 10]


def test_x32_headline_batch_ramps_then_holds_the_target():
    n = 4096
    in_len, out_len = [300 * 1024] * n, [MiB] * n
    c = _plan(False, in_len, out_len, first_byte=4)                    # X_32 order-0 streams
    b = _chunk_bytes(c, in_len, out_len)
    assert 10 <= len(b) <= 20
    assert b[0] < b[1] < b[2] < b[3]                                   # 1/8, 1/4, 1/2, then full-size chunks
    assert (b[3:-1] <= 384 * MiB).all() and (b[3:-1] > 300 * MiB).all()


def test_four_way_streams_get_larger_chunks():
    n = 4096
    in_len, out_len = [300 * 1024] * n, [MiB] * n
    x32 = _chunk_bytes(_plan(False, in_len, out_len, first_byte=4), in_len, out_len)
    way4 = _chunk_bytes(_plan(False, in_len, out_len, first_byte=0), in_len, out_len)
    r4x8 = _chunk_bytes(_plan(False, in_len, out_len, first_byte=0, method=[1] * n), in_len, out_len)
    assert len(way4) == 3 and len(r4x8) == 3                           # capped at a third of the batch
    assert way4.max() > 4 * x32.max()
    # small 4-way blocks (CRAM-sized): their serial time is short again, so the default target applies
    n2 = 40000
    small = _plan(False, [30 * 1024] * n2, [100 * 1024] * n2, first_byte=1)
    assert len(small) - 1 >= 8


def test_encode_uses_the_order_word():
    n = 2048
    raw, cap = [MiB] * n, [hb.rans_compress_bound_4x16(MiB, 0)] * n
    x32 = _plan(True, raw, cap, order=[hb.RANS_ORDER_X32] * n)
    way4 = _plan(True, raw, cap, order=[1] * n)
    legacy = _plan(True, raw, cap, order=[hb.ORDER_RANS4x8] * n)
    assert 3 <= len(way4) <= 4 and len(legacy) == len(way4) and len(x32) > len(way4)   # 2-3 chunks vs many
    # one 4-way block among X_32 ones sets the pace of its chunk, hence of the call
    mixed = _plan(True, raw, cap, order=[hb.RANS_ORDER_X32] * (n - 1) + [0])
    assert len(mixed) == len(way4)


def test_ragged_sizes_tile_exactly():
    rng = np.random.default_rng(1)
    out_len = rng.integers(1, 3 * MiB, 900)
    in_len = (out_len * rng.uniform(0.05, 1.0, 900)).astype(np.int64) + 1
    c = _plan(False, in_len, out_len, first_byte=rng.integers(0, 256, 900).astype(np.uint8))
    b = _chunk_bytes(c, in_len, out_len)
    assert b.sum() == int(in_len.sum() + out_len.sum())


def test_encode_chunks_are_equal_and_few():
    """Encode is bound by the host->device stream, which starts at once: no small leading chunks, and as few chunks
    as the kernels' latency floor allows (each chunk costs one round of latency-bound kernels)."""
    n = 4096
    raw, cap = [MiB] * n, [hb.rans_compress_bound_4x16(MiB, 0)] * n
    x32 = _chunk_bytes(_plan(True, raw, cap, order=[hb.RANS_ORDER_X32] * n), raw, cap)
    assert 4 <= len(x32) <= 16
    assert x32[:-1].max() - x32[:-1].min() <= 3 * MiB                  # equal but for block granularity
    dec = _chunk_bytes(_plan(False, [300 * 1024] * n, raw, first_byte=4), [300 * 1024] * n, raw)
    assert len(dec) > len(x32) and dec[0] < dec[3] / 4                 # the decoder still ramps up
    way4 = _chunk_bytes(_plan(True, raw, cap, order=[0] * n), raw, cap)
    assert 2 <= len(way4) <= 4 and len(way4) < len(x32)               # a 4-way round of kernels is ~5 x longer
