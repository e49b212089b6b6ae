"""The chunked host pipeline (several chunks per call, pinned host memory) in both copy modes:
full duplex (three staging slots reused round-robin) and half duplex (one slot per chunk, inputs first)."""
import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import synth

pytestmark = pytest.mark.gpu
N = 1 << 20


@pytest.mark.parametrize("full", [True, False])
def test_multi_chunk_pinned_roundtrip(full, oracle):
    ctx = hb.Context(0)
    ctx.set_copy_duplex(full)
    distinct, nblk = 6, 120
    raw = [synth.qual_block(50 + i, N).tobytes() for i in range(distinct)]
    for order in (4, 1):
        want = [oracle.compress(d, order) for d in raw]
        # ---- encode: raw blocks in pinned memory -> capacity-strided output regions
        pin_raw = hb.PinnedArray(nblk * N)
        for i in range(nblk):
            pin_raw.array[i * N:(i + 1) * N] = np.frombuffer(raw[i % distinct], np.uint8)
        cap = (hb.rans_compress_bound_4x16(N, order) + 15) // 16 * 16
        pin_c = hb.PinnedArray(nblk * cap)
        r_off = np.arange(nblk, dtype=np.uint64) * N
        c_off = np.arange(nblk, dtype=np.uint64) * cap
        r_len = np.full(nblk, N, np.uint32)
        status = np.zeros(nblk, np.int32)
        for rep in range(2):                                  # the second call predicts stream lengths from the first
            c_len = np.full(nblk, cap, np.uint32)
            pin_c.array[:] = 0
            ctx.compress_batch_host(nblk, pin_raw.array, r_off, r_len, pin_c.array, c_off, c_len, status,
                                    np.full(nblk, order, np.int32))
            assert (status == 0).all()
            for i in range(nblk):
                w = want[i % distinct]
                assert int(c_len[i]) == len(w) and bytes(pin_c.array[i * cap: i * cap + len(w)]) == w, (order, rep, i)
        # ---- decode them back into pinned memory
        pin_out = hb.PinnedArray(nblk * N)
        out_len = np.full(nblk, N, np.uint32)
        ctx.uncompress_batch_host(nblk, pin_c.array, c_off, c_len, pin_out.array, r_off, out_len, status)
        assert (status == 0).all() and (out_len == N).all()
        assert np.array_equal(pin_out.array, pin_raw.array)
    ctx.close()
