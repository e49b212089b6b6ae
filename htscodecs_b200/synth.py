"""Seeded synthetic block generators.

Plain numpy; used by bench.py and tests/ so that the GPU box can regenerate every input from a
seed instead of shipping data.
"""
import numpy as np

ILLUMINA_8BIN = (np.array([2, 6, 15, 22, 27, 33, 37, 40]) + 33).astype(np.uint8)
ILLUMINA_P = np.array([.01, .02, .03, .05, .08, .15, .36, .30])


def _sticky(rng, n, symbols, p, stay):
    """First-order sticky chain: keep the previous symbol with probability `stay`, else redraw
    from `p` (so `p` is also the stationary distribution)."""
    if n == 0:
        return np.zeros(0, np.uint8)
    draws = rng.choice(len(symbols), size=n, p=p / p.sum())
    keep = rng.random(n) < stay
    keep[0] = False
    src = np.where(keep, 0, np.arange(n))
    np.maximum.accumulate(src, out=src)
    return symbols[draws[src]]


def qual_block(block, n=1 << 20, stay=0.6):
    """Config 2/3: Illumina 8-bin qualities, numpy default_rng(1234 + block)."""
    rng = np.random.default_rng(1234 + block)
    return _sticky(rng, n, ILLUMINA_8BIN, ILLUMINA_P, stay)


def acgt_block(block, n=1 << 20):
    """Config 4(i): i.i.d. ACGT with p = .3/.2/.2/.3."""
    rng = np.random.default_rng(4321 + block)
    return np.frombuffer(b"ACGT", np.uint8)[rng.choice(4, size=n, p=[.3, .2, .2, .3])]


def tag_block(block, n=1 << 20, nsym=6, mean_run=30):
    """Config 4(ii): low-entropy tag data, <= nsym symbols in geometric runs."""
    rng = np.random.default_rng(9876 + block)
    syms = rng.choice(256, size=nsym, replace=False).astype(np.uint8)
    nruns = max(1, int(n / mean_run * 1.3) + 16)
    lens = rng.geometric(1.0 / mean_run, size=nruns)
    vals = syms[rng.integers(0, nsym, size=nruns)]
    out = np.repeat(vals, lens)
    while out.size < n:
        out = np.concatenate([out, out])
    return np.ascontiguousarray(out[:n])


def u32_block(block, n=1 << 20):
    """Config 4(iii): little-endian u32 array of slowly increasing integers (STRIPE food)."""
    rng = np.random.default_rng(555 + block)
    a = np.cumsum(rng.integers(0, 50, size=(n + 3) // 4)).astype("<u4")
    return np.frombuffer(a.tobytes(), np.uint8)[:n].copy()


def wide_block(block, n=1 << 20, nsym=45):
    """~45-symbol unbinned-quality-like data (order-1 tables too big for shared memory)."""
    rng = np.random.default_rng(777 + block)
    p = rng.dirichlet(np.ones(nsym) * 0.5)
    syms = (np.arange(nsym) + 33).astype(np.uint8)
    return _sticky(rng, n, syms, p, 0.4)


def random_block(block, n=1 << 20):
    rng = np.random.default_rng(31337 + block)
    return rng.integers(0, 256, size=n, dtype=np.uint8)


GENERATORS = {
    "qual": qual_block, "acgt": acgt_block, "tag": tag_block,
    "u32": u32_block, "wide": wide_block, "random": random_block,
}


def mixed_corpus(nblk, seed=5, size=1 << 20, ragged=True):
    """Config 5: a mixed-flag CRAM block corpus.  Returns a list of
    (generator, block_seed, n_bytes, flags, method) with method 0 = rANS 4x16 (flags = the
    reference's `order` argument) and 1 = legacy rANS 4x8 (flags = order 0/1).  Shares follow the
    survey's example mix: 40 % o0, 30 % o1, 10 % X_32 o0, 10 % X_32 o1, 5 % PACK/RLE/STRIPE
    variants, 5 % rANS 4x8."""
    rng = np.random.default_rng(seed)
    out = []
    variants = [("acgt", 0x80), ("acgt", 0x81), ("tag", 0x40), ("tag", 0xc1), ("tag", 0xc5), ("u32", 0x408),
                ("u32", 0x409), ("tag", 0x3c9), ("acgt", 0x84), ("u32", 0x40c)]
    for i in range(nblk):
        u = rng.random()
        n = int(size if (not ragged or rng.random() < 0.7) else rng.integers(1, size + 1))
        if u < 0.40:
            gen, flags, method = ("qual", "wide")[int(rng.random() < 0.2)], 0, 0
        elif u < 0.70:
            gen, flags, method = ("qual", "wide")[int(rng.random() < 0.2)], 1, 0
        elif u < 0.80:
            gen, flags, method = "qual", 4, 0
        elif u < 0.90:
            gen, flags, method = "qual", 5, 0
        elif u < 0.95:
            gen, flags = variants[int(rng.integers(0, len(variants)))]
            method = 0
        else:
            gen, flags, method = "qual", int(rng.integers(0, 2)), 1
        out.append((gen, i, n, flags, method))
    return out
