"""Pins the CPU oracle (oracle/hts_oracle.c) to the reference.  CPU only.

* the reference's own golden streams (tests/dat/r4x16, r4x8 -> tests/golden/): decode AND encode
  must match byte for byte (mirrors tests/rans4x16.test:14-28 and rans4x8.test:9-27);
* the reference's varint known answers (tests/varint_test.c:146-154);
* ref_vectors.json: answers of the unmodified reference on the cases its tests do not pin;
* when oracle/_ref/libref.so is present, a randomised differential run against it.
"""
import hashlib
import os

import numpy as np
import pytest

from vectors import small_inputs, large_cases, ALL_FLAGS
from htscodecs_b200 import synth

GOLD16 = {
    "q4": [0, 1, 64, 65, 128, 129, 192, 193, 8, 9],
    "q8": [0, 1, 64, 65, 128, 129, 192, 193],
    "q40+dir": [0, 1, 8, 9],
    "qvar": [0, 1],
}


def _src(golden_dir, name):
    with open(os.path.join(golden_dir, "src", name + ".bin"), "rb") as f:
        return f.read()


@pytest.mark.parametrize("name,flag", [(n, f) for n, fl in GOLD16.items() for f in fl])
def test_golden_4x16(oracle, golden_dir, name, flag):
    data = _src(golden_dir, name)
    with open(os.path.join(golden_dir, "r4x16", f"{name}.{flag}"), "rb") as f:
        gold = f.read()
    assert oracle.uncompress(gold, len(data)) == data          # rans4x16.test:27-28
    assert oracle.compress(data, flag) == gold                 # fresh encode == golden


def test_config1_known_answer(oracle, golden_dir):
    """BASELINE config 1: q40+dir order-0 -> 50247 bytes, md5 77e4167d..."""
    data = _src(golden_dir, "q40+dir")
    assert len(data) == 100000 and hashlib.md5(data).hexdigest() == "ea2e88c7a117c3989203f6987058d548"
    c = oracle.compress(data, 0)
    assert len(c) == 50247 and hashlib.md5(c).hexdigest() == "77e4167d5c157de511d7e14573796d88"


@pytest.mark.parametrize("name", list(GOLD16))
@pytest.mark.parametrize("order", [0, 1])
def test_golden_4x8_decode(oracle, golden_dir, name, order):
    data = _src(golden_dir, name)
    with open(os.path.join(golden_dir, "r4x8", f"{name}.{order}"), "rb") as f:
        gold = f.read()
    assert oracle.uncompress_4x8(gold) == data                 # rans4x8.test:24-25


def test_varint_known_answers(oracle):
    # tests/varint_test.c:146-154
    assert oracle.var_put(0x80) == bytes([0x81, 0x00])
    assert oracle.var_put(0x1234) == bytes([0xa4, 0x34])
    assert oracle.var_put(0xffffffff) == bytes([0x8f, 0xff, 0xff, 0xff, 0x7f])
    assert oracle.var_put(0) == b"\0" and oracle.var_put(0x7f) == b"\x7f"


def test_bound_known_answers(oracle):
    # SURVEY 8a, measured on the reference
    assert oracle.bound(1048576, 0) == 1101802
    assert oracle.bound(1048576, 1) == 1299952
    assert oracle.bound(1048576, 64) == 1300728


def test_ref_vectors_small(oracle, ref_vectors):
    inputs = dict(small_inputs())
    n = 0
    for v in ref_vectors["small"]:
        data = inputs[v["name"]]
        assert hashlib.md5(data).hexdigest() == v["in_md5"], "generator drifted: " + v["name"]
        c = oracle.compress(data, v["flags"])
        assert c is not None and len(c) == v["clen"], (v["name"], hex(v["flags"]))
        assert hashlib.md5(c).hexdigest() == v["out_md5"], (v["name"], hex(v["flags"]))
        if v["out"] is not None:
            assert c.hex() == v["out"]
        d = oracle.uncompress(c, len(data))
        if v["ref_decodes"]:
            assert d == data, (v["name"], hex(v["flags"]))
        else:
            assert d is None, (v["name"], hex(v["flags"]))     # e.g. the :948 reject quirk
        n += 1
    assert n > 1000


def test_ref_vectors_4x8(oracle, reflib, ref_vectors):
    """4x8 streams are produced by the reference encoder (out of scope for us) so this needs libref."""
    inputs = dict(small_inputs())
    for v in ref_vectors["small_4x8"]:
        data = inputs[v["name"]]
        c8 = reflib.compress_4x8(data, v["order"])
        assert hashlib.md5(c8).hexdigest() == v["out_md5"]
        assert oracle.uncompress_4x8(c8) == data, v["name"]


@pytest.mark.parametrize("case", large_cases(), ids=lambda c: f"{c[0]}-{c[4]:#x}")
def test_ref_vectors_large(oracle, ref_vectors, case):
    name, gen, block, n, flags = case
    v = [x for x in ref_vectors["large"] if x["name"] == name and x["flags"] == flags and "codec" not in x][0]
    data = synth.GENERATORS[gen](block, n).tobytes()
    assert hashlib.md5(data).hexdigest() == v["in_md5"], "generator drifted"
    c = oracle.compress(data, flags)
    assert len(c) == v["clen"] and hashlib.md5(c).hexdigest() == v["out_md5"]
    assert oracle.uncompress(c, n) == data


def test_x32_roundtrip_and_layout(oracle):
    """X_32 is the N=32 generalisation (parity unpinned): round trips, flag bit kept, and the
    stream differs from the 4-way one only by the 28 extra states for order-0 tables."""
    for gen in ("qual", "wide", "tag", "acgt", "u32"):
        data = synth.GENERATORS[gen](5, 70001).tobytes()
        for f in (4, 5, 0x44, 0x45, 0x84, 0x85, 0xc4, 0xc5, 0x0c, 0x0d, 0xcd):
            c = oracle.compress(data, f)
            assert c is not None
            if not (c[0] & 0x20) and not (f & 8):
                assert c[0] & 4
            assert oracle.uncompress(c, len(data)) == data, (gen, hex(f))
    data = synth.qual_block(0, 1 << 16).tobytes()
    c4, c32 = oracle.compress(data, 0), oracle.compress(data, 4)
    # same table bytes: the streams share everything up to the first state word
    tab = len(c4) - 16
    k = next(i for i in range(min(len(c4), len(c32))) if c4[i] != c32[i])
    assert k == 0 and c4[1:8] == c32[1:8]
    assert abs(len(c32) - len(c4) - 112) < 64 and tab > 0


def test_small_x32(oracle):
    for name, data in small_inputs():
        for f in (4, 5, 0x45, 0x85, 0xc5, 0x0d):
            c = oracle.compress(data, f)
            assert c is not None and oracle.uncompress(c, len(data)) == data, (name, hex(f))


def test_differential_vs_reference(oracle, reflib):
    rng = np.random.default_rng(99)
    kinds = ["qual", "wide", "tag", "acgt", "u32", "random"]
    for it in range(120):
        gen = kinds[it % len(kinds)]
        n = int(rng.integers(21, 60000))
        data = synth.GENERATORS[gen](1000 + it, n).tobytes()
        for f in ALL_FLAGS:
            if (f & 8) and (f >> 8) > n:
                continue
            a, b = reflib.compress(data, f), oracle.compress(data, f)
            assert a == b, (gen, n, hex(f))
            assert oracle.uncompress(a, n) == reflib.uncompress(a, n)
        for order in (0, 1):
            c8 = reflib.compress_4x8(data, order)
            assert oracle.uncompress_4x8(c8) == data


# ------------------------------------------------------------------ rANS 4x8 encoder
@pytest.mark.parametrize("name", list(GOLD16))
@pytest.mark.parametrize("order", [0, 1])
def test_golden_4x8_encode(oracle, golden_dir, name, order):
    """The oracle's 4x8 encoder reproduces the reference's pre-compressed r4x8 files."""
    data = open(os.path.join(golden_dir, "src", name + ".bin"), "rb").read()
    gold = open(os.path.join(golden_dir, "r4x8", f"{name}.{order}"), "rb").read()
    assert oracle.compress_4x8(data, order) == gold


def test_4x8_encode_differential_vs_reference(oracle, reflib):
    from vectors import small_inputs
    from htscodecs_b200 import synth
    cases = [(n, d) for n, d in small_inputs() if len(d) >= 1]
    cases += [(g, synth.GENERATORS[g](2, 70001).tobytes()) for g in ("qual", "wide", "tag", "random", "acgt", "u32")]
    for name, data in cases:
        for order in (0, 1):
            assert oracle.compress_4x8(data, order) == reflib.compress_4x8(data, order), (name, order)
    assert oracle.compress_4x8(b"", 0) is None


def test_x32_pins(oracle):
    """X_32 regression pins (tests/golden/x32_pins.json, made by make_x32_pins.py from THIS oracle: the
    reference has no X_32, so they freeze our definition rather than prove parity)."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "x32_pins.json")) as f:
        pins = json.load(f)
    assert len(pins) >= 90
    for p in pins:
        data = synth.GENERATORS[p["gen"]](p["block"], p["n"]).tobytes()
        assert hashlib.md5(data).hexdigest() == p["in_md5"], "generator drifted"
        c = oracle.compress(data, p["flags"])
        assert len(c) == p["clen"] and hashlib.md5(c).hexdigest() == p["out_md5"], (p["gen"], p["n"], hex(p["flags"]))
        assert c[:24].hex() == p["head"]
