"""Container glue on the GPU path: framed files (`[u32 clen][stream]...`, the format the reference's test programs
write, tests/rANS_static4x16pr_test.c:261-296) decoded in ONE batched call via hts_b200_frames_scan, and written
back via hts_b200_frames_write."""
import os
import struct

import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_framed_buffer_roundtrip(oracle):
    ctx = hb.Context(0)
    raw = [synth.GENERATORS[g](i, 20000 + 1001 * i).tobytes() for i, g in enumerate(["qual", "tag", "acgt", "u32", "wide", "qual"])]
    orders = [0, 0xc1, 0x80, 9, 1, 5]
    streams, st = ctx.compress_many(raw, orders)
    assert (st == 0).all() and streams == [oracle.compress(d, f) for d, f in zip(raw, orders)]
    lens = np.array([len(s) for s in streams], np.uint32)
    off = np.zeros(len(streams), np.uint64); off[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    framed = hb.frames_write(np.frombuffer(b"".join(streams), np.uint8), off, lens)
    assert framed == b"".join(struct.pack("=I", len(s)) + s for s in streams)
    buf = np.frombuffer(framed, np.uint8)
    in_off, in_len, out_off, out_len, total = hb.frames_scan(buf, out_align=16)
    out = np.zeros(total + 16, np.uint8)
    got_len = out_len.copy()
    status = np.zeros(len(raw), np.int32)
    ctx.uncompress_batch_host(len(raw), buf, in_off, in_len, out, out_off, got_len, status)
    assert (status == 0).all()
    for i, d in enumerate(raw):
        assert bytes(out[int(out_off[i]): int(out_off[i]) + int(got_len[i])]) == d
    ctx.close()


def test_file_written_by_the_reference_program_decodes_in_one_call(tmp_path):
    """The reference's own CLI (compiled unmodified against this library, oracle/Makefile `dropin`) writes a framed
    file block by block; the batched decoder takes the whole file."""
    from test_gpu_dropin_programs import _run, _read
    data = synth.qual_block(9, 3 * 1043156 + 12345).tobytes()          # four blocks of the program's BLK_SIZE
    src, dst = tmp_path / "in.bin", tmp_path / "out.r4x16"
    src.write_bytes(data)
    _run("rans4x16pr_b200", "-o1", str(src), str(dst))                 # no -r: `[u32 clen][stream]` per block
    comp = _read(str(dst))
    ctx = hb.Context(0)
    buf = np.frombuffer(comp, np.uint8)
    in_off, in_len, out_off, out_len, total = hb.frames_scan(buf)
    assert len(in_len) == 4 and total == len(data)
    out = np.zeros(total, np.uint8)
    status = np.zeros(4, np.int32)
    ctx.uncompress_batch_host(4, buf, in_off, in_len, out, out_off, out_len, status)
    assert (status == 0).all() and out.tobytes() == data
    ctx.close()
