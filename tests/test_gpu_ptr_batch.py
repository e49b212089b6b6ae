"""The pointer-array batch calls (rans4x16_uncompress_batch / rans4x16_compress_batch, include/htscodecs_b200.h):
one pageable buffer per block, as an htslib-style caller holds them.  Several chunks per call, so the per-chunk scatter
into the caller's buffers runs while the pipeline is still moving the later chunks."""
import ctypes as C

import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import synth

pytestmark = pytest.mark.gpu
N = 300 * 1024


def _ptrs(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


@pytest.mark.parametrize("orders", [(0, 1), (4, 5), (0x40 | 0, 0x80 | 1, hb.ORDER_RANS4x8 | 1)])
def test_pointer_array_roundtrip_matches_oracle(orders, oracle):
    lib = hb.load_library()
    ctx = hb.Context(0)
    distinct, nblk = 6, 301                                   # ~115 MiB per direction: four or more chunks; 301 is not a multiple of 8
    gens = [synth.qual_block, synth.acgt_block, synth.tag_block]
    raw = [gens[i % 3](70 + i, N).tobytes() for i in range(distinct)]
    order = [orders[i % len(orders)] for i in range(nblk)]
    want = {}
    for i in range(nblk):
        key = (i % distinct, order[i])
        if key not in want:
            want[key] = (oracle.compress_4x8(raw[key[0]], order[i] & 1) if order[i] & hb.ORDER_RANS4x8
                         else oracle.compress(raw[key[0]], order[i]))
    # ---- encode
    ins = [np.frombuffer(raw[i % distinct], np.uint8).copy() for i in range(nblk)]
    caps = np.array([lib.hts_b200_compress_bound_4x8(N) if o & hb.ORDER_RANS4x8 else hb.rans_compress_bound_4x16(N, o)
                     for o in order], np.uint32)
    outs = [np.full(int(c) + 32, 0xEE, np.uint8) for c in caps]
    isz = np.full(nblk, N, np.uint32)
    osz = caps.copy()
    st = np.full(nblk, 99, np.int32)
    ordv = np.array(order, np.int32)
    rc = lib.rans4x16_compress_batch(ctx.h, nblk, C.cast(_ptrs(ins), C.c_void_p), isz.ctypes.data,
                                     C.cast(_ptrs(outs), C.c_void_p), osz.ctypes.data, ordv.ctypes.data, st.ctypes.data)
    assert rc == 0, ctx.last_error()
    assert (st == 0).all(), st[st != 0][:8]
    for i in range(nblk):
        w = want[(i % distinct, order[i])]
        assert int(osz[i]) == len(w) and bytes(outs[i][: len(w)]) == w, (i, order[i])
        assert (outs[i][int(caps[i]):] == 0xEE).all()         # nothing beyond the block's capacity
    # ---- decode what was just written, method 1 for the 4x8 streams
    if any(o & hb.ORDER_RANS4x8 for o in order):
        # the pointer-array decoder is the 4x16 entry point: check those blocks only
        keep = [i for i in range(nblk) if not order[i] & hb.ORDER_RANS4x8]
    else:
        keep = list(range(nblk))
    cins = [outs[i][: int(osz[i])].copy() for i in keep]
    douts = [np.full(N + 16, 0xEE, np.uint8) for _ in keep]
    m = len(keep)
    cisz = np.array([a.size for a in cins], np.uint32)
    dosz = np.full(m, N, np.uint32)
    dst = np.full(m, 99, np.int32)
    rc = lib.rans4x16_uncompress_batch(ctx.h, m, C.cast(_ptrs(cins), C.c_void_p), cisz.ctypes.data,
                                       C.cast(_ptrs(douts), C.c_void_p), dosz.ctypes.data, dst.ctypes.data)
    assert rc == 0, ctx.last_error()
    assert (dst == 0).all() and (dosz == N).all()
    for k, i in enumerate(keep):
        assert bytes(douts[k][:N]) == raw[i % distinct], i
        assert (douts[k][N:] == 0xEE).all()
    ctx.close()
