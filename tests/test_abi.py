"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/htscodecs_b200.h
declares (no compute calls here -- there is no GPU in the authoring container)."""
import ctypes
import os
import re

import pytest

import htscodecs_b200 as hb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "htscodecs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(\w+)\s*\(", src))
    return {n for n in names if n.startswith(("rans", "hts_b200_"))}


def test_header_and_binding_agree():
    assert _declared() == set(hb.EXPORTS)


def test_library_exports_every_declared_symbol():
    from htscodecs_b200 import build
    build.build()
    lib = ctypes.CDLL(hb.LIB_PATH)
    for name in sorted(_declared()):
        assert getattr(lib, name) is not None, name


def test_bound_matches_reference_numbers():
    # pure host arithmetic (reference rANS_static4x16pr.c:360-372); SURVEY 8a values
    assert hb.rans_compress_bound_4x16(1048576, 0) == 1101802
    assert hb.rans_compress_bound_4x16(1048576, 1) == 1299952
    assert hb.rans_compress_bound_4x16(1048576, 64) == 1300728


def test_bound_matches_oracle(oracle):
    for n in (0, 1, 20, 21, 1000, 65536, 1 << 20, (1 << 24) + 3):
        for order in (0, 1, 4, 5, 64, 65, 128, 193, 8, 9, 0xc9, 0x208, 0x2009):
            assert hb.rans_compress_bound_4x16(n, order) == oracle.bound(n, order)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        return
    try:
        hb.Context(0)
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context() must fail without a GPU")
    assert hb.rans_uncompress_4x16(b"\x00\x05hello") is None


def test_product_does_not_touch_the_oracle():
    """The product path may not import / link / execute anything under oracle/."""
    pkg = os.path.join(ROOT, "htscodecs_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("no CPU fallback", ""), os.path.join(dirpath, fn)


def test_reference_test_programs_link_against_the_library():
    """oracle/Makefile `dropin` compiles the reference's own test programs against
    libhtscodecs_b200.so; every rans_* symbol they import must be one the library defines."""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref")
    for prog in ("rans4x16pr_b200", "rans4x8_b200"):
        exe = os.path.join(ref, prog)
        if not os.path.exists(exe):
            pytest.skip("drop-in programs not built (reference tree absent at build time)")
        und = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
        wanted = {ln.split()[-1] for ln in und.splitlines() if " rans_" in ln}
        assert wanted, prog
        lib = subprocess.run(["nm", "-D", "--defined-only", hb.LIB_PATH], capture_output=True, text=True,
                             check=True).stdout
        have = {ln.split()[-1] for ln in lib.splitlines()}
        assert wanted <= have, wanted - have


def test_multi_device_call_fails_loudly_without_a_gpu():
    """No CPU fallback behind the multi-device entry either: without an sm_100 device it returns an error."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        return
    n = 2
    buf = np.zeros(64, np.uint8)
    off = np.zeros(n, np.uint64)
    ln = np.full(n, 8, np.uint32)
    st = np.zeros(n, np.int32)
    with pytest.raises(RuntimeError) as e:
        hb.uncompress_batch_host_multi([0], n, buf, off, ln, buf, off, ln.copy(), st)
    assert "no CPU fallback" in str(e.value)
    assert hb.multi_last_stats() is not None and hb.multi_launch_count() == 0
    lib = hb.load_library()
    assert lib.hts_b200_scratch_bytes(None) == 0
