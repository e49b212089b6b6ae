"""GPU parity tests for the legacy rANS 4x8 ENCODER: streams produced by the
CUDA kernels through the C ABI must equal the reference's golden files and the CPU oracle's output
byte for byte, and decode back (on the GPU) to the source."""
import os

import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import synth
from vectors import small_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = hb.Context(0)
    yield c
    c.close()


def test_golden_files_are_reproduced(golden_dir):
    for name in ("q4", "q8", "q40+dir", "qvar"):
        data = open(os.path.join(golden_dir, "src", name + ".bin"), "rb").read()
        for order in (0, 1):
            gold = open(os.path.join(golden_dir, "r4x8", f"{name}.{order}"), "rb").read()
            assert hb.rans_compress(data, order) == gold, (name, order)
    assert hb.rans_compress(b"", 0) is None


def test_vectors_vs_oracle_batched(ctx, oracle):
    cases = [(n, d) for n, d in small_inputs() if len(d) >= 1]
    for gen in ("qual", "wide", "tag", "random", "acgt", "u32"):
        cases.append((gen + "300k", synth.GENERATORS[gen](3, 300007).tobytes()))
        cases.append((gen + "1M", synth.GENERATORS[gen](1, 1 << 20).tobytes()))
    blocks, orders, want, labels = [], [], [], []
    for name, data in cases:
        for order in (0, 1):
            blocks.append(data); orders.append(order | hb.ORDER_RANS4x8)
            want.append(oracle.compress_4x8(data, order)); labels.append((name, order))
    got, status = ctx.compress_many(blocks, orders)
    bad = [(l, int(s), None if g is None else len(g), len(w)) for l, g, s, w in zip(labels, got, status, want) if g != w]
    assert not bad, bad[:8]
    # and back: one batched GPU decode of everything just produced
    out, st = ctx.uncompress_many(got, [len(b) for b in blocks], [1] * len(blocks))
    assert (st == 0).all() and out == blocks


def test_mixed_codecs_in_one_batch(ctx, oracle):
    """4x16 and 4x8 blocks side by side in one compress call."""
    data = [synth.qual_block(i, 50000 + 13 * i).tobytes() for i in range(12)]
    orders = [0, 1, 4, 5, hb.ORDER_RANS4x8, 1 | hb.ORDER_RANS4x8] * 2
    got, status = ctx.compress_many(data, orders)
    assert (status == 0).all()
    for d, o, g in zip(data, orders, got):
        want = oracle.compress_4x8(d, o & 1) if o & hb.ORDER_RANS4x8 else oracle.compress(d, o)
        assert g == want, hex(o)
