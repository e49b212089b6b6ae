"""Host-side multi-GPU logic on CPU: partition of the block list and the gather of
per-block sizes/status, run as two `gloo` processes.  The codec worker here is the CPU oracle
(tests may use it; the product never does) so that the gathered sizes can be checked for real."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from htscodecs_b200 import shard, synth  # noqa: E402


def test_partition_balanced_and_contiguous():
    rng = np.random.default_rng(1)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 3, 8, 100, 4096):
            w = rng.integers(1, 1 << 20, size=n)
            r = shard.partition_blocks(w, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
            for a, b in zip(r[:-1], r[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
            if n >= 8 * world:
                tot = w.sum()
                for lo, hi in r:
                    assert abs(w[lo:hi].sum() - tot / world) <= w.max()


def test_partition_equal_blocks_is_even():
    r = shard.partition_blocks([1 << 20] * 4096, 8)
    assert [hi - lo for lo, hi in r] == [512] * 8


def test_partition_skewed():
    w = [100] + [1] * 10
    r = shard.partition_blocks(w, 2)
    assert r == [(0, 1), (1, 11)]


def test_exclusive_offsets():
    off = shard.exclusive_offsets([5, 16, 1], align=16)
    assert off.tolist() == [0, 16, 32]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle()
        sizes = [3000 + 517 * i for i in range(11)]
        flags = [0, 1, 4, 5, 0x40, 0x80, 0xc1, 9, 0, 1, 4]
        blocks = [synth.qual_block(i, sizes[i]).tobytes() for i in range(11)]

        def encode(lo, hi):
            comps = [o.compress(blocks[i], flags[i]) for i in range(lo, hi)]
            encode.comps = comps
            return [len(c) for c in comps], [0] * (hi - lo)

        ranges, clen, status = shard.run_sharded(encode, sizes, rank, world, dist)
        # every rank now knows every compressed size -> the packed-arena offsets of the whole job
        off = shard.exclusive_offsets(clen)
        lo, hi = ranges[rank]
        ok = (status == 0).all() and all(len(c) == clen[lo + k] for k, c in enumerate(encode.comps))
        # decode leg: this rank decodes its own range; sizes gathered again
        def decode(lo, hi):
            outs = [o.uncompress(encode.comps[i - lo], sizes[i]) for i in range(lo, hi)]
            good = [0 if outs[k] == blocks[lo + k] else -1 for k in range(hi - lo)]
            return [len(x) for x in outs], good

        _, ulen, st2 = shard.run_sharded(decode, sizes, rank, world, dist)
        ok = ok and ulen.tolist() == sizes and (st2 == 0).all()
        q.put((rank, bool(ok), ranges, clen.tolist(), off.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    # both ranks hold identical, complete results
    assert res[0][2:] == res[1][2:]
    ranges, clen, off = res[0][2], res[0][3], res[0][4]
    assert ranges[0][1] == ranges[1][0] and ranges[1][1] == 11
    assert all(c > 0 for c in clen) and off[0] == 0 and off[-1] == sum(clen[:-1])
