"""the name tokeniser's back end tries up to nine rANS 4x16 orders per byte column and
keeps the smallest (`compress()`, tokenise_name3.c:1246-1299).  rans4x16_compress_best_batch does that
selection for many columns in one batched device pass; winner and bytes must match the loop run with the
CPU oracle."""
import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import synth

pytestmark = pytest.mark.gpu

# the reference's level table (tokenise_name3.c:1254-1260); 193+8 = PACK | RLE | ORDER1 | STRIPE
LEVELS = {1: [0, 128], 3: [0, 192 + 8], 5: [0, 128, 193 + 8], 7: [0, 1, 129, 65, 193, 193 + 8],
          9: [0, 1, 128, 129, 64, 65, 192, 193, 193 + 8]}


def _columns():
    """Byte columns like the tokeniser's: token types (few symbols, long runs), digit strings, deltas,
    fixed-width integers; sizes from 0 to a few thousand, multiples of 4 and not."""
    rng = np.random.default_rng(7)
    cols = []
    for i in range(60):
        n = int(rng.integers(0, 6000))
        kind = i % 6
        if kind == 0:
            c = synth.tag_block(i, max(n, 1), nsym=4, mean_run=40)[:n]
        elif kind == 1:
            c = (rng.integers(0, 10, n) + 48).astype(np.uint8)
        elif kind == 2:
            c = synth.u32_block(i, max(n // 4 * 4, 4))[: n // 4 * 4]
        elif kind == 3:
            c = np.minimum(rng.geometric(0.4, n), 255).astype(np.uint8)
        elif kind == 4:
            c = synth.acgt_block(i, max(n, 1))[:n]
        else:
            c = synth.qual_block(i, max(n, 1))[:n]
        cols.append(np.ascontiguousarray(c).tobytes())
    cols += [b"", b"A", b"AAAA", bytes(range(256)) * 3]
    return cols


def _expected(oracle, data, methods):
    best_sz, best, stream = None, None, None
    for m in methods:
        if len(data) % 4 != 0 and (m & 8):                               # :1269
            continue
        c = oracle.compress(data, m)
        if best_sz is None or len(c) < best_sz:                          # strictly smaller wins (:1280)
            best_sz, best, stream = len(c), m, c
    return best, stream


@pytest.mark.parametrize("level", [3, 9])
def test_best_of_matches_reference_selection(level, oracle):
    ctx = hb.Context(0)
    cols = _columns()
    methods = LEVELS[level]
    got, best, status = ctx.compress_best_many(cols, methods)
    assert (status == 0).all(), status
    for i, d in enumerate(cols):
        want_best, want = _expected(oracle, d, methods)
        assert best[i] == want_best, (i, len(d), best[i], want_best)
        assert got[i] == want, (i, len(d))
    # and the winners decode
    out, st = ctx.uncompress_many(got, [len(d) for d in cols])
    assert (st == 0).all() and out == cols
    ctx.close()


def test_best_of_large_columns_and_capacity(oracle):
    """1 MiB columns (several passes' worth of candidates stay on the device; only winners return)
    and the capacity contract."""
    ctx = hb.Context(0)
    cols = [synth.qual_block(1, 1 << 20).tobytes(), synth.tag_block(2, 1 << 20).tobytes(), synth.acgt_block(3, (1 << 20) - 1).tobytes()]
    got, best, status = ctx.compress_best_many(cols, LEVELS[9])
    assert (status == 0).all()
    for i, d in enumerate(cols):
        want_best, want = _expected(oracle, d, LEVELS[9])
        assert best[i] == want_best and got[i] == want
    ctx.close()
