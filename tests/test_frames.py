"""CPU-only: container glue (`[u32 clen][stream]...`, reference tests/rANS_static4x16pr_test.c:261-296)
and the block partition of the multi-device calls -- host arithmetic of the C library, no GPU needed.
Streams come from the CPU oracle (test infrastructure)."""
import struct

import numpy as np

import htscodecs_b200 as hb
from htscodecs_b200 import shard, synth


def _framed(streams):
    return b"".join(struct.pack("=I", len(s)) + s for s in streams)


def test_frames_scan_matches_the_framing(oracle):
    raw = [synth.qual_block(i, 1000 + 777 * i).tobytes() for i in range(7)]
    orders = [0, 1, 4, 5, 0x40, 0x81, 9]
    comp = [oracle.compress(d, f) for d, f in zip(raw, orders)]
    buf = _framed(comp)
    got = hb.frames_scan(buf, out_align=16)
    assert got is not None
    in_off, in_len, out_off, out_len, total = got
    pos = 0
    want_out = 0
    for i, c in enumerate(comp):
        assert int(in_off[i]) == pos + 4 and int(in_len[i]) == len(c)
        assert bytes(buf[int(in_off[i]): int(in_off[i]) + len(c)]) == c
        assert int(out_len[i]) == len(raw[i]) and int(out_off[i]) == want_out
        want_out += (len(raw[i]) + 15) // 16 * 16
        pos += 4 + len(c)
    assert total == want_out


def test_frames_scan_4x8_and_empty(oracle):
    raw = [synth.qual_block(3, 5000).tobytes(), synth.acgt_block(1, 300).tobytes()]
    comp = [oracle.compress_4x8(d, i & 1) for i, d in enumerate(raw)]
    in_off, in_len, out_off, out_len, total = hb.frames_scan(_framed(comp), method=hb.RANS4x8)
    assert out_len.tolist() == [5000, 300] and out_off.tolist() == [0, 5000] and total == 5300
    assert hb.frames_scan(b"")[4] == 0 and len(hb.frames_scan(b"")[0]) == 0


def test_frames_scan_rejects_truncation_and_nosz(oracle):
    c = oracle.compress(b"A" * 100 + b"CGT" * 50, 0)
    good = _framed([c, c])
    assert hb.frames_scan(good) is not None
    assert hb.frames_scan(good[:-1]) is None                      # stream cut short
    assert hb.frames_scan(good + b"\x01\x00") is None             # dangling partial length field
    nosz = oracle.compress(b"A" * 100 + b"CGT" * 50, hb.RANS_ORDER_NOSZ)
    assert hb.frames_scan(_framed([nosz])) is None                # no stored size: cannot be batched blind


def test_frames_write_is_the_inverse(oracle):
    raw = [synth.tag_block(i, 4000).tobytes() for i in range(5)]
    comp = [oracle.compress(d, 0x40) for d in raw]
    lens = np.array([len(c) for c in comp], np.uint32)
    off = np.zeros(5, np.uint64)
    off[1:] = np.cumsum((lens[:-1].astype(np.uint64) + 63) // 64 * 64)
    arena = np.zeros(int(off[-1]) + int(lens[-1]), np.uint8)
    for i, c in enumerate(comp):
        arena[int(off[i]): int(off[i]) + len(c)] = np.frombuffer(c, np.uint8)
    assert hb.frames_write(arena, off, lens) == _framed(comp)
    status = np.array([0, -1, 0, 0, -2], np.int32)                 # failed blocks are left out
    assert hb.frames_write(arena, off, lens, status) == _framed([comp[0], comp[2], comp[3]])
    # and the result scans back to the same streams
    in_off, in_len, _, out_len, _ = hb.frames_scan(hb.frames_write(arena, off, lens))
    assert in_len.tolist() == lens.tolist() and out_len.tolist() == [4000] * 5


def test_partition_is_the_c_rule():
    rng = np.random.default_rng(7)
    for world in (1, 2, 3, 8):
        for n in (0, 1, 5, 64, 1000):
            w = rng.integers(0, 1 << 20, size=n).astype(np.uint32)
            r = shard.partition_blocks(w, world)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            # each cut is the boundary nearest to its share of the total weight
            csum = np.concatenate([[0], np.cumsum(w.astype(np.int64))])
            tot = int(csum[-1])
            for k in range(1, world):
                c = r[k][0]
                if tot == 0:
                    continue
                tgt = tot * k / world
                best = min(abs(int(csum[j]) - tgt) for j in range(r[k - 1][0], n + 1))
                assert abs(int(csum[c]) - tgt) <= best + 1e-6
    assert shard.partition_blocks([0, 0, 0, 0], 2) == [(0, 2), (2, 4)]
