#!/usr/bin/env python
"""Regression pins for X_32 streams: tests/golden/x32_pins.json.

The mounted reference (htscodecs v1.1) has no 32-way interleave, so these are NOT reference outputs
("parity unpinned"): they are the outputs of oracle/hts_oracle.c -- the N-way
restatement that is pinned to the reference at N = 4 -- frozen so that the X_32 stream definition
cannot drift unnoticed between rounds.  Inputs are regenerated from seeds (htscodecs_b200.synth).

    python tests/golden/make_x32_pins.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from oracle_lib import Oracle  # noqa: E402
from htscodecs_b200 import synth  # noqa: E402

CASES = [(gen, block, n, f)
         for gen, block, n in (("qual", 0, 1 << 20), ("qual", 3, 300001), ("wide", 0, 1 << 18), ("tag", 1, 100003),
                               ("acgt", 0, 1 << 18), ("u32", 0, 1 << 18), ("random", 0, 1 << 16), ("qual", 9, 31),
                               ("qual", 9, 32), ("qual", 9, 33), ("qual", 9, 1000))
         for f in (4, 5, 0x44, 0x45, 0x84, 0x85, 0xc5, 0x0c, 0x0d)]


def main():
    o = Oracle()
    pins = []
    for gen, block, n, f in CASES:
        data = synth.GENERATORS[gen](block, n).tobytes()
        c = o.compress(data, f)
        assert o.uncompress(c, n) == data
        pins.append({"gen": gen, "block": block, "n": n, "flags": f, "in_md5": hashlib.md5(data).hexdigest(),
                     "clen": len(c), "out_md5": hashlib.md5(c).hexdigest(), "head": c[:24].hex()})
    with open(os.path.join(HERE, "x32_pins.json"), "w") as fh:
        json.dump(pins, fh, indent=0)
    print(len(pins), "pins")


if __name__ == "__main__":
    main()
