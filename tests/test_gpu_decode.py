"""GPU parity tests for the decode path: every stream is produced by the CPU oracle (pinned to the
reference), decoded by the CUDA kernels THROUGH THE C ABI, and compared byte for byte."""
import os

import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import synth
from vectors import small_inputs, large_cases, ALL_FLAGS

pytestmark = pytest.mark.gpu

GOLD16 = {
    "q4": [0, 1, 64, 65, 128, 129, 192, 193, 8, 9],
    "q8": [0, 1, 64, 65, 128, 129, 192, 193],
    "q40+dir": [0, 1, 8, 9],
    "qvar": [0, 1],
}
X32_FLAGS = [4, 5, 0x44, 0x45, 0x84, 0x85, 0xc4, 0xc5, 0x0c, 0x0d, 0xcd]


@pytest.fixture(scope="module")
def ctx():
    c = hb.Context(0)
    yield c
    c.close()


def _src(golden_dir, name):
    with open(os.path.join(golden_dir, "src", name + ".bin"), "rb") as f:
        return f.read()


def _report(tag, cases, out, status, expect):
    bad = []
    for (label, _), o, s, e in zip(cases, out, status, expect):
        if o != e:
            first = None
            if o is not None and e is not None:
                n = min(len(o), len(e))
                first = next((i for i in range(n) if o[i] != e[i]), n)
            bad.append((label, int(s), None if o is None else len(o), None if e is None else len(e), first))
    assert not bad, f"{tag}: {len(bad)} of {len(cases)} differ; first few (label, status, got, want, first diff): {bad[:8]}"


def test_golden_streams_dropin(golden_dir):
    """The reference's own pre-compressed files through rans_uncompress_4x16 / rans_uncompress."""
    for name, flags in GOLD16.items():
        data = _src(golden_dir, name)
        for f in flags:
            comp = open(os.path.join(golden_dir, "r4x16", f"{name}.{f}"), "rb").read()
            assert hb.rans_uncompress_4x16(comp) == data, (name, f)
        for order in (0, 1):
            comp = open(os.path.join(golden_dir, "r4x8", f"{name}.{order}"), "rb").read()
            assert hb.rans_uncompress(comp) == data, (name, order)


def test_golden_streams_batched(ctx, golden_dir):
    streams, expect, methods, cases = [], [], [], []
    for name, flags in GOLD16.items():
        data = _src(golden_dir, name)
        for f in flags:
            streams.append(open(os.path.join(golden_dir, "r4x16", f"{name}.{f}"), "rb").read())
            expect.append(data); methods.append(0); cases.append((f"{name}.{f}", None))
        for order in (0, 1):
            streams.append(open(os.path.join(golden_dir, "r4x8", f"{name}.{order}"), "rb").read())
            expect.append(data); methods.append(1); cases.append((f"4x8/{name}.{order}", None))
    out, status = ctx.uncompress_many(streams, [len(e) for e in expect], methods)
    _report("golden", cases, out, status, expect)


def test_small_vectors(ctx, oracle):
    """Every size/flag quirk of vectors.small_inputs, 4-way and X_32; error cases must also agree
    (e.g. the reference rejects its own all-256-symbol order-1 stream, …4x16pr.c:948)."""
    streams, sizes, expect, cases = [], [], [], []
    for name, data in small_inputs():
        for f in ALL_FLAGS + X32_FLAGS:
            if (f & 8) and len(data) > 20 and (f >> 8) > len(data):
                continue
            c = oracle.compress(data, f)
            assert c is not None
            streams.append(c); sizes.append(len(data)); cases.append((f"{name}/{f:#x}", None))
            expect.append(oracle.uncompress(c, len(data)))
    out, status = ctx.uncompress_many(streams, sizes)
    _report("small", cases, out, status, expect)


@pytest.mark.parametrize("x32", [0, 4])
def test_large_vectors(ctx, oracle, x32):
    streams, sizes, expect, cases = [], [], [], []
    for name, gen, block, n, flags in large_cases():
        data = synth.GENERATORS[gen](block, n).tobytes()
        c = oracle.compress(data, flags | x32)
        streams.append(c); sizes.append(n); expect.append(data); cases.append((f"{name}/{flags | x32:#x}", None))
    out, status = ctx.uncompress_many(streams, sizes)
    _report("large", cases, out, status, expect)


def test_4x8_vectors(ctx, reflib, oracle):
    streams, expect, cases = [], [], []
    for name, data in small_inputs():
        if len(data) < 1:
            continue
        for order in (0, 1):
            c = reflib.compress_4x8(data, order)
            if c is None:
                continue
            e = oracle.uncompress_4x8(c)
            streams.append(c); expect.append(e); cases.append((f"{name}/o{order}", None))
    for gen in ("qual", "wide", "tag", "random"):
        data = synth.GENERATORS[gen](3, 300007).tobytes()
        for order in (0, 1):
            c = reflib.compress_4x8(data, order)
            streams.append(c); expect.append(data); cases.append((f"{gen}300k/o{order}", None))
    sizes = [len(e) if e is not None else hb.peek_size(s, 1) for s, e in zip(streams, expect)]
    out, status = ctx.uncompress_many(streams, sizes, [1] * len(streams))
    _report("4x8", cases, out, status, expect)


def test_device_resident_batch(ctx, oracle):
    """hts_b200_uncompress_batch_dev with torch-owned device memory, 64 x 1 MiB X_32 blocks (config 2 shape)."""
    import torch
    nblk, n = 64, 1 << 20
    blocks = [synth.qual_block(i, n).tobytes() for i in range(8)]
    comps = [oracle.compress(b, 4) for b in blocks]
    in_len = np.array([len(comps[i % 8]) for i in range(nblk)], np.uint32)
    in_off = np.zeros(nblk, np.uint64)
    in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64))           # dense, hence odd/even alignments
    buf = np.concatenate([np.frombuffer(comps[i % 8], np.uint8) for i in range(nblk)])
    d_in = torch.from_numpy(buf).cuda()
    d_in_off = torch.from_numpy(in_off.view(np.int64)).cuda()
    d_in_len = torch.from_numpy(in_len.view(np.int32)).cuda()
    out_off = (np.arange(nblk, dtype=np.uint64) * n)
    d_out = torch.zeros(nblk * n, dtype=torch.uint8, device="cuda")
    d_out_off = torch.from_numpy(out_off.view(np.int64)).cuda()
    d_out_len = torch.full((nblk,), n, dtype=torch.int32, device="cuda")
    d_status = torch.full((nblk,), -99, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.uncompress_batch_dev(nblk, d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_len, d_status)
    assert (d_status.cpu().numpy() == 0).all()
    assert (d_out_len.cpu().numpy() == n).all()
    res = d_out.cpu().numpy()
    for i in range(nblk):
        assert res[i * n:(i + 1) * n].tobytes() == blocks[i % 8], i


def test_size_contract(ctx, oracle):
    data = synth.qual_block(1, 5000).tobytes()
    c = oracle.compress(data, 0)
    # capacity too small -> error status, like the reference's NULL (…4x16pr.c:1464)
    out, status = ctx.uncompress_many([c], [len(data) - 1])
    assert out[0] is None and status[0] != 0
    # larger capacity is fine for non-striped streams
    out, status = ctx.uncompress_many([c], [len(data) + 100])
    assert out[0] == data
    # striped streams need the exact size through the drop-in call (…4x16pr.c:1379)
    cs = oracle.compress(data, 9)
    assert hb.rans_uncompress_to_4x16(cs, len(data)) == data
    assert hb.rans_uncompress_to_4x16(cs, len(data) + 1) is None
    # X_NOSZ needs the caller's size
    cn = oracle.compress(data, 0x10)
    assert hb.rans_uncompress_4x16(cn) is None
    assert hb.rans_uncompress_to_4x16(cn, len(data)) == data
    # garbage / truncated input fails cleanly
    assert hb.rans_uncompress_4x16(b"") is None
    out, status = ctx.uncompress_many([c[: len(c) // 2], b"\x00", c, b"\x08\x05"], [len(data)] * 4)
    assert out[2] == data
    assert out[1] == b""            # the reference also accepts this: flags 0, no size, no body -> 0 bytes
    assert out[3] is None           # truncated stripe header


def test_fuzz_no_crash(ctx, oracle):
    """Bit-flipped streams must not fault the device (outputs are unspecified, statuses may vary)."""
    rng = np.random.default_rng(5)
    streams, sizes = [], []
    for it in range(300):
        gen = ["qual", "wide", "tag", "acgt", "u32"][it % 5]
        data = synth.GENERATORS[gen](it, 3000 + it).tobytes()
        f = int(rng.choice(ALL_FLAGS + X32_FLAGS))
        c = bytearray(oracle.compress(data, f))
        for _ in range(int(rng.integers(1, 4))):
            c[int(rng.integers(0, len(c)))] ^= 1 << int(rng.integers(0, 8))
        streams.append(bytes(c)); sizes.append(len(data))
    ctx.uncompress_many(streams, sizes)
    # the context must still work afterwards
    data = synth.qual_block(2, 10000).tobytes()
    out, status = ctx.uncompress_many([oracle.compress(data, 5)], [len(data)])
    assert out[0] == data


def test_fuzz_large_batch_all_codecs(ctx, oracle):
    """Damaged streams of every codec (4x16 families, X_32, rANS 4x8, compressed order-1 tables) in
    batches large enough for the high-occupancy kernel variants, through both batch entry points.
    The device must not fault (tools/fuzz.py is the long-running form of this test)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from fuzz import damage
    rng = np.random.default_rng(11)
    base = []
    for i in range(60):
        gen = ["qual", "wide", "tag", "acgt", "u32", "random"][i % 6]
        n = int(rng.integers(1, 5000))
        d = synth.GENERATORS[gen](i, n).tobytes()
        base.append((oracle.compress(d, (ALL_FLAGS + X32_FLAGS)[i % len(ALL_FLAGS + X32_FLAGS)]), n, 0))
        base.append((oracle.compress_4x8(d, i & 1), n, 1))
    d2 = bytes(rng.permutation(np.arange(256).repeat(8)).astype(np.uint8))
    base.append((oracle.compress(d2, 1), len(d2), 0))
    base.append((oracle.compress(d2, 5), len(d2), 0))
    for it in range(2):
        streams, sizes, methods = [], [], []
        for _ in range(5600):
            c, n, m = base[int(rng.integers(0, len(base)))]
            streams.append(damage(rng, c)); sizes.append(n); methods.append(m)
        (ctx.uncompress_many_dev if it else ctx.uncompress_many)(streams, sizes, methods)
        data = synth.qual_block(2, 10000).tobytes()
        out, status = ctx.uncompress_many([oracle.compress(data, 5)], [len(data)])
        assert status[0] == 0 and out[0] == data
