#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference checkout (run in the authoring container only).

    python tests/golden/make_golden.py [/root/reference]

1. src/<name>.bin      = `cut -f1 tests/dat/<name> | tr -d '\\n'`  (what tests/rans4x16.test:11 feeds the codec)
2. r4x16/*, r4x8/*     = the reference's own pre-compressed golden streams, copied verbatim (DATA, not source)
3. ref_vectors.json    = outputs of the unmodified reference (oracle/_ref/libref.so) on seeded synthetic
                         inputs for the cases the reference's tests do not pin: tiny sizes,
                         PACK/RLE quirks, STRIPE with N != 4, CAT, 1 MiB blocks.  Inputs are regenerated from
                         seeds (tests/vectors.py); each case stores length + md5 of the reference stream (and the
                         stream itself in hex when it is <= 40 bytes).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from oracle_lib import RefLib  # noqa: E402
from vectors import small_cases, large_cases  # noqa: E402


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    dat = os.path.join(ref_root, "tests", "dat")
    os.makedirs(os.path.join(HERE, "src"), exist_ok=True)
    for name in ("q4", "q8", "q40+dir", "qvar"):
        with open(os.path.join(dat, name), "rb") as f:
            col = b"".join(line.split(b"\t")[0].rstrip(b"\n") for line in f)
        with open(os.path.join(HERE, "src", name + ".bin"), "wb") as f:
            f.write(col)
        print(name, len(col), hashlib.md5(col).hexdigest())
    for sub in ("r4x16", "r4x8"):
        dst = os.path.join(HERE, sub)
        os.makedirs(dst, exist_ok=True)
        for fn in sorted(os.listdir(os.path.join(dat, sub))):
            shutil.copyfile(os.path.join(dat, sub, fn), os.path.join(dst, fn))

    ref = RefLib()
    vec = {"small": [], "large": [], "small_4x8": []}
    for name, data, flags in small_cases():
        c = ref.compress(data, flags)
        d = ref.uncompress(c, len(data)) if c is not None else None
        vec["small"].append({"name": name, "flags": flags, "in_md5": hashlib.md5(data).hexdigest(),
                             "clen": None if c is None else len(c),
                             "out_md5": None if c is None else hashlib.md5(c).hexdigest(),
                             "out": c.hex() if c is not None and len(c) <= 40 else None,
                             "ref_decodes": d == data})
        if len(data) >= 1 and flags in (0, 1):
            c8 = ref.compress_4x8(data, flags)
            if c8 is not None and ref.uncompress_4x8(c8) == data:
                vec["small_4x8"].append({"name": name, "order": flags, "in_md5": hashlib.md5(data).hexdigest(),
                                         "clen": len(c8), "out_md5": hashlib.md5(c8).hexdigest()})
    for name, gen, block, n, flags in large_cases():
        from htscodecs_b200 import synth
        data = synth.GENERATORS[gen](block, n).tobytes()
        c = ref.compress(data, flags)
        d = ref.uncompress(c, len(data))
        assert d == data, (name, flags)
        vec["large"].append({"name": name, "gen": gen, "block": block, "n": n, "flags": flags,
                             "in_md5": hashlib.md5(data).hexdigest(),
                             "clen": len(c), "out_md5": hashlib.md5(c).hexdigest()})
        if flags in (0, 1):
            c8 = ref.compress_4x8(data, flags)
            assert ref.uncompress_4x8(c8) == data
            vec["large"].append({"name": name + "/4x8", "gen": gen, "block": block, "n": n, "flags": flags,
                                 "codec": "4x8", "in_md5": hashlib.md5(data).hexdigest(),
                                 "clen": len(c8), "out_md5": hashlib.md5(c8).hexdigest()})
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as f:
        json.dump(vec, f, indent=0)
    print("small", len(vec["small"]), "small_4x8", len(vec["small_4x8"]), "large", len(vec["large"]))


if __name__ == "__main__":
    main()
