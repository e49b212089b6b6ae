"""The drop-in claim, exercised the way the reference tests itself: the reference's OWN test
programs (`tests/rANS_static4x16pr_test.c`, `tests/rANS_static_test.c`; compiled UNMODIFIED by
`oracle/Makefile` target `dropin`, linked against libhtscodecs_b200.so instead of
rANS_static4x16pr.c / rANS_static.c) run the reference's own test scripts
(`tests/rans4x16.test:11-30`, `tests/rans4x8.test:11-28`) over the reference's golden corpus.

The scripts only ask for a round trip and for the pre-compressed files to decode; we also ask for
the compressed bytes to equal the golden files (the encoder is byte-identical, so they must)."""
import os
import struct
import subprocess

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(HERE, "golden")
FILES = ("q4", "q8", "q40+dir", "qvar")


def _env():
    env = dict(os.environ)
    extra = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cuda_runtime as cr                      # the wheel that ships libcudart.so.12
        extra.insert(0, os.path.join(list(cr.__path__)[0], "lib"))
    except ImportError:
        pass
    env["LD_LIBRARY_PATH"] = ":".join(extra + [env.get("LD_LIBRARY_PATH", "")])
    return env


def _run(prog, *args):
    exe = os.path.join(REFDIR, prog)
    if not os.path.exists(exe):
        if os.path.isdir("/root/reference/tests"):                         # authoring container: build them now
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "dropin"], check=True, capture_output=True)
        else:
            pytest.skip(f"{exe} was not built (`make -C oracle dropin` needs the reference tree) and did not travel")
    r = subprocess.run([exe, *args], env=_env(), capture_output=True, timeout=300)
    assert r.returncode == 0, (prog, args, r.stderr[-400:])
    return r


def _read(p):
    with open(p, "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", FILES)
def test_rans4x16_test_script(name, tmp_path):
    src = os.path.join(GOLD, "src", name + ".bin")
    data = _read(src)
    comp, uncomp = str(tmp_path / "r4x16.comp"), str(tmp_path / "r4x16.uncomp")
    ran = 0
    for o in (0, 1, 64, 65, 128, 129, 192, 193, 8, 9):
        gold = os.path.join(GOLD, "r4x16", f"{name}.{o}")
        if not os.path.exists(gold):
            continue
        _run("rans4x16pr_b200", "-r", f"-o{o}", src, comp)                 # rans4x16.test:22
        assert _read(comp) == _read(gold), (name, o)
        _run("rans4x16pr_b200", "-r", "-d", comp, uncomp)                  # :24
        assert _read(uncomp) == data
        _run("rans4x16pr_b200", "-r", "-d", gold, uncomp)                  # :28 pre-compressed data
        assert _read(uncomp) == data
        ran += 1
    assert ran >= 2


@pytest.mark.parametrize("name", FILES)
def test_rans4x8_test_script(name, tmp_path):
    src = os.path.join(GOLD, "src", name + ".bin")
    data = _read(src)
    comp, uncomp = str(tmp_path / "r4x8.comp"), str(tmp_path / "r4x8.uncomp")
    for o in (0, 1):
        gold = os.path.join(GOLD, "r4x8", f"{name}.{o}")
        _run("rans4x8_b200", "-r", f"-o{o}", src, comp)                    # rans4x8.test:20
        assert _read(comp) == _read(gold), (name, o)
        _run("rans4x8_b200", "-r", "-d", comp, uncomp)
        assert _read(uncomp) == data
        _run("rans4x8_b200", "-r", "-d", gold, uncomp)
        assert _read(uncomp) == data


def test_framed_and_timing_modes(tmp_path, oracle):
    """The programs' other modes: `[u32 size][stream]` framing of BLK_SIZE = 1039*251*4 byte blocks
    (rANS_static4x16pr_test.c:261-296; the striped order `-o9.4` syntax of :100-105) and the `-t`
    loop that calls rans_compress_to_4x16 / rans_uncompress_to_4x16 on caller buffers
    (:191-207) and prints "Mismatch" on a bad round trip."""
    from htscodecs_b200 import synth
    blk = 1039 * 251 * 4                                                   # rANS_static4x16pr_test.c:48
    data = b"".join(synth.qual_block(i, blk).tobytes() for i in range(3))[: 2 * blk + 12345]
    src, comp, uncomp = str(tmp_path / "in"), str(tmp_path / "comp"), str(tmp_path / "out")
    with open(src, "wb") as f:
        f.write(data)
    for o, flags in (("1", 1), ("193", 193), ("9.4", 9 | 4 << 8)):
        _run("rans4x16pr_b200", f"-o{o}", src, comp)
        blob, pos, k = _read(comp), 0, 0
        while pos < len(blob):                                             # each frame equals the oracle's stream
            (sz,) = struct.unpack_from("<I", blob, pos)
            want = oracle.compress(data[k * blk: (k + 1) * blk], flags)
            assert blob[pos + 4: pos + 4 + sz] == want, (o, k)
            pos += 4 + sz
            k += 1
        assert k == 3
        _run("rans4x16pr_b200", "-d", comp, uncomp)
        assert _read(uncomp) == data
    r = _run("rans4x16pr_b200", "-t", "-o1", src)
    assert b"Mismatch" not in r.stderr and b"MB/s enc" in r.stderr
    r = _run("rans4x8_b200", "-t", "-o1", src)
    assert b"Mismatch" not in r.stderr
