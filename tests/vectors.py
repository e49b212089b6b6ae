"""Case lists shared by tests/golden/make_golden.py (which asks the reference for the answers)
and the parity tests (which replay them against the oracle and the CUDA path)."""
import numpy as np

FLAGS_BASIC = [0, 1, 64, 65, 128, 129, 192, 193, 8, 9]
FLAGS_EXTRA = [0xc9, 0x208, 0x3c9, 0x809, 0x20, 0x48, 0x88, 0x1009, 0x2009]
ALL_FLAGS = FLAGS_BASIC + FLAGS_EXTRA


def _runs(rng, n, nsym, mean):
    syms = rng.integers(0, 256, nsym, dtype=np.uint8)
    out = []
    tot = 0
    while tot < n:
        ln = int(rng.geometric(1.0 / mean))
        out.append(np.full(ln, syms[rng.integers(0, nsym)], np.uint8))
        tot += ln
    return np.concatenate(out)[:n].tobytes() if n else b""


def small_inputs():
    """(name, bytes) -- deterministic small inputs covering the reference's size quirks."""
    rng = np.random.default_rng(20240517)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    cases = []
    for n in (0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 19, 20, 21, 22, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 1000, 1023, 4097):
        cases.append((f"acgt{n}", acgt[rng.integers(0, 4, n)].tobytes()))
        cases.append((f"rand{n}", rng.integers(0, 256, n, dtype=np.uint8).tobytes()))
    for n in (30, 500, 5000):
        cases.append((f"const{n}", bytes([65]) * n))
        cases.append((f"two{n}", bytes(rng.choice([0, 255], n).astype(np.uint8))))
        cases.append((f"runs{n}", _runs(rng, n, 3, 12)))
        cases.append((f"runs6_{n}", _runs(rng, n, 6, 30)))
        cases.append((f"sym17_{n}", bytes(rng.integers(100, 117, n, dtype=np.uint8))))
        cases.append((f"sym16_{n}", bytes(rng.integers(100, 116, n, dtype=np.uint8))))
        cases.append((f"sym5_{n}", bytes(rng.integers(0, 5, n, dtype=np.uint8))))
        cases.append((f"u32_{n}", np.cumsum(rng.integers(0, 9, (n + 3) // 4)).astype("<u4").tobytes()[:n]))
    cases.append(("all256x20", bytes(range(256)) * 20))          # reference decoder rejects o1 (:948)
    cases.append(("all256x3", bytes(range(256)) * 3))
    cases.append(("all256rand", bytes(rng.permutation(np.arange(256).repeat(12)).astype(np.uint8))))
    cases.append(("skew", bytes(rng.choice(256, 6000, p=np.r_[0.97, np.full(255, 0.03 / 255)]).astype(np.uint8))))
    cases.append(("zeros_then", bytes(3000) + bytes(rng.integers(0, 256, 500, dtype=np.uint8))))
    return cases


def small_cases():
    for name, data in small_inputs():
        for f in ALL_FLAGS:
            # the reference's transpose loop underflows when N > in_size (rANS_static4x16pr.c:1173)
            if (f & 8) and len(data) > 20 and (f >> 8) > len(data):
                continue
            yield name, data, f


def large_cases():
    """(name, generator, block, n, flags) -- regenerated from seeds via htscodecs_b200.synth."""
    M = 1 << 20
    out = []
    for f in (0, 1):
        out.append(("qual1M", "qual", 0, M, f))
        out.append(("qual1M_b7", "qual", 7, M, f))
        out.append(("qualBLK", "qual", 3, 1039 * 251 * 4, f))      # the reference harness block size
        out.append(("wide1M", "wide", 0, M, f))
        out.append(("random256k", "random", 0, 1 << 18, f))
        out.append(("qual300k", "qual", 1, 300001, f))
    for f in (128, 129):
        out.append(("acgt1M", "acgt", 0, M, f))
    for f in (64, 65, 192, 193):
        out.append(("tag1M", "tag", 0, M, f))
        out.append(("tag100k", "tag", 1, 100003, f))
    for f in (8, 9, 0xc9, 0x208, 0x809):
        out.append(("u32_1M", "u32", 0, M, f))
    for f in (0xc9, 0x3c9):
        out.append(("tag1M_stripe", "tag", 2, M, f))
    return out
