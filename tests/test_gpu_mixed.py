"""Config 5 (BASELINE.json configs[4], scaled down): a mixed-flag corpus -- rANS 4x16 o0/o1, X_32,
PACK/RLE/STRIPE variants and legacy rANS 4x8 -- in ONE batched call per direction.  Encoded
streams must equal the CPU oracle's byte for byte; decoded blocks must equal the source."""
import hashlib

import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import shard, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = hb.Context(0)
    yield c
    c.close()


def _corpus(nblk, size):
    spec = synth.mixed_corpus(nblk, seed=5, size=size)
    blocks = [synth.GENERATORS[g](b, n).tobytes() for g, b, n, _, _ in spec]
    return spec, blocks


def test_mixed_corpus_parity(ctx, oracle, reflib):
    spec, blocks = _corpus(160, 200_000)
    i16 = [i for i, s in enumerate(spec) if s[4] == 0]
    i8 = [i for i, s in enumerate(spec) if s[4] == 1]
    assert i16 and i8
    # ---- encode every 4x16 block in one call; compare with the oracle
    comp, status = ctx.compress_many([blocks[i] for i in i16], [spec[i][3] for i in i16])
    assert (status == 0).all()
    for k, i in enumerate(i16):
        assert comp[k] == oracle.compress(blocks[i], spec[i][3]), (i, spec[i])
    # ---- one mixed decode batch: our 4x16 streams + reference-made 4x8 streams
    streams = [None] * len(spec)
    for k, i in enumerate(i16):
        streams[i] = comp[k]
    for i in i8:
        streams[i] = reflib.compress_4x8(blocks[i], spec[i][3])
    out, st = ctx.uncompress_many(streams, [len(b) for b in blocks], [s[4] for s in spec])
    assert (st == 0).all(), st
    assert out == blocks


def test_mixed_corpus_full_size_roundtrip(ctx):
    """1 MiB blocks (ragged tail sizes included): GPU encode -> GPU decode, checked through a
    checksum of checksums so the test stays cheap at full block size."""
    spec, blocks = _corpus(96, 1 << 20)
    i16 = [i for i, s in enumerate(spec) if s[4] == 0]
    comp, status = ctx.compress_many([blocks[i] for i in i16], [spec[i][3] for i in i16])
    assert (status == 0).all()
    out, st = ctx.uncompress_many(comp, [len(blocks[i]) for i in i16])
    assert (st == 0).all()
    want = hashlib.sha256(b"".join(hashlib.sha256(blocks[i]).digest() for i in i16)).hexdigest()
    got = hashlib.sha256(b"".join(hashlib.sha256(o).digest() for o in out)).hexdigest()
    assert got == want
    # and the partition the multi-GPU path would use keeps every rank within one block of the mean
    w = [len(blocks[i]) for i in i16]
    for lo, hi in shard.partition_blocks(w, 8):
        assert abs(sum(w[lo:hi]) - sum(w) / 8) <= max(w)


def test_big_blocks(ctx, oracle):
    """Blocks far larger than anything tiled on chip (counter folds, ring refills, 32-bit offsets):
    a 40 MiB and a ragged 9 MiB block through every entropy codec, byte-exact against the oracle."""
    big = np.concatenate([synth.qual_block(100 + i, 1 << 20) for i in range(40)]).tobytes()
    odd = big[5: 5 + 9 * (1 << 20) + 12345]
    blocks, orders = [], []
    for data in (big, odd):
        for f in (0, 1, 4, 5, hb.ORDER_RANS4x8, hb.ORDER_RANS4x8 | 1):
            blocks.append(data); orders.append(f)
    comp, status = ctx.compress_many(blocks, orders)
    assert (status == 0).all()
    for d, f, c in zip(blocks, orders, comp):
        want = oracle.compress_4x8(d, f & 1) if f & hb.ORDER_RANS4x8 else oracle.compress(d, f)
        assert c == want, hex(f)
    out, st = ctx.uncompress_many(comp, [len(b) for b in blocks], [1 if f & hb.ORDER_RANS4x8 else 0 for f in orders])
    assert (st == 0).all() and out == blocks


def test_dropin_calls_from_many_host_threads(oracle):
    """The reference is re-entrant (one block per thread, rANS_static4x16pr.c:853-858); the drop-in
    symbols keep a context per host thread, so concurrent callers must not disturb each other."""
    from concurrent.futures import ThreadPoolExecutor

    def work(t):
        bad = 0
        for i in range(12):
            d = synth.qual_block(1000 * t + i, 30000 + 977 * i).tobytes()
            f = (0, 1, 4, 5, 0x40, 0x81)[i % 6]
            c = hb.rans_compress_4x16(d, f)
            bad += c != oracle.compress(d, f)
            bad += hb.rans_uncompress_4x16(c) != d
            c8 = hb.rans_compress(d, i & 1)
            bad += c8 != oracle.compress_4x8(d, i & 1)
            bad += hb.rans_uncompress(c8) != d
        return bad

    with ThreadPoolExecutor(max_workers=6) as ex:
        assert sum(ex.map(work, range(6))) == 0


def test_large_batch_routes_to_compact_order0_kernels(ctx, oracle, reflib):
    """More than 5000 blocks in one device-resident call: small-alphabet 4-way / 4x8 order-0 streams
    and <= 9-symbol 4-way order-1 streams are decoded by the high-occupancy kernel variants (code
    paths no other test reaches)."""
    import torch
    kinds = [("qual", 0, 0), ("acgt", 0, 0), ("wide", 0, 0), ("tag", 0x40, 0), ("random", 0, 0), ("qual", 0, 1),
             ("wide", 0, 1), ("qual", 1, 0), ("qual", 4, 0), ("acgt", 1, 0), ("wide", 1, 0), ("tag", 0x41, 0),
             ("qual", 1, 1)]
    uniq, comp, meth = [], [], []
    for i, (gen, f, m) in enumerate(kinds):
        for n in (4096 + 37 * i, 3, 700):
            d = synth.GENERATORS[gen](i, n).tobytes()
            uniq.append(d); meth.append(m)
            comp.append(reflib.compress_4x8(d, f) if m else oracle.compress(d, f))
    nblk = 5400
    idx = [i % len(uniq) for i in range(nblk)]
    in_len = np.array([len(comp[i]) for i in idx], np.uint32)
    in_off = np.zeros(nblk, np.uint64); in_off[1:] = np.cumsum((in_len[:-1].astype(np.uint64) + 15) // 16 * 16)
    out_cap = np.array([len(uniq[i]) for i in idx], np.uint32)
    out_off = np.zeros(nblk, np.uint64); out_off[1:] = np.cumsum((out_cap[:-1].astype(np.uint64) + 15) // 16 * 16)
    ib = np.zeros(int(in_off[-1] + in_len[-1]) + 16, np.uint8)
    for k, i in enumerate(idx):
        ib[int(in_off[k]): int(in_off[k]) + len(comp[i])] = np.frombuffer(comp[i], np.uint8)
    d_in = torch.from_numpy(ib).cuda()
    d_out = torch.zeros(int(out_off[-1] + out_cap[-1]) + 16, dtype=torch.uint8, device="cuda")
    d_len = torch.from_numpy(out_cap.view(np.int32).copy()).cuda()
    d_st = torch.zeros(nblk, dtype=torch.int32, device="cuda")
    ctx.uncompress_batch_dev(nblk, d_in, torch.from_numpy(in_off.view(np.int64)).cuda(),
                             torch.from_numpy(in_len.view(np.int32)).cuda(), d_out,
                             torch.from_numpy(out_off.view(np.int64)).cuda(), d_len, d_st,
                             torch.from_numpy(np.array([meth[i] for i in idx], np.uint8)).cuda())
    assert int((d_st != 0).sum()) == 0
    ob = d_out.cpu().numpy()
    for k in list(range(0, 60)) + list(range(nblk - 60, nblk)):
        assert bytes(ob[int(out_off[k]): int(out_off[k]) + int(out_cap[k])]) == uniq[idx[k]], k
