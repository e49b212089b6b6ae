"""hts_b200_{un,}compress_batch_host_multi: one host-buffer batch cut over the GPUs of this process (one thread +
context per device, copy phases coordinated across devices).  Runs on however many GPUs the box shows: with one
GPU the partition logic is still exercised through devices=[0]."""
import numpy as np
import pytest

import htscodecs_b200 as hb
from htscodecs_b200 import shard, synth

pytestmark = pytest.mark.gpu
N = 1 << 18


def _devices():
    import torch
    return list(range(min(torch.cuda.device_count(), 8)))


@pytest.mark.parametrize("phased", [True, False])
def test_multi_roundtrip_matches_the_checker(phased, oracle):
    devs = _devices()
    hb.multi_set_phased(phased)
    rng = np.random.default_rng(11)
    distinct, nblk = 10, 300
    gens = ["qual", "tag", "acgt", "wide", "u32"]
    orders1 = [0, 1, 4, 5, 0x40, 0x81, 0xc5, 9, 0x44, 0]
    sizes = [int(N if i % 3 else rng.integers(1000, N)) // 4 * 4 for i in range(distinct)]
    raw = [synth.GENERATORS[gens[i % 5]](i, sizes[i]).tobytes() for i in range(distinct)]
    want = [oracle.compress(d, f) for d, f in zip(raw, orders1)]
    idx = [i % distinct for i in range(nblk)]
    # ---- encode over all devices
    r_len = np.array([len(raw[k]) for k in idx], np.uint32)
    r_off = np.zeros(nblk, np.uint64); r_off[1:] = np.cumsum(r_len[:-1].astype(np.uint64))
    pin_raw = hb.PinnedArray(int(r_len.astype(np.uint64).sum()))
    for i, k in enumerate(idx):
        pin_raw.array[int(r_off[i]): int(r_off[i]) + len(raw[k])] = np.frombuffer(raw[k], np.uint8)
    orders = np.array([orders1[k] for k in idx], np.int32)
    caps = np.array([hb.rans_compress_bound_4x16(int(r_len[i]), int(orders[i])) for i in range(nblk)], np.uint32)
    c_off = np.zeros(nblk, np.uint64); c_off[1:] = np.cumsum((caps[:-1].astype(np.uint64) + 15) // 16 * 16)
    pin_c = hb.PinnedArray(int(c_off[-1]) + int(caps[-1]))
    c_len = caps.copy()
    status = np.full(nblk, 77, np.int32)
    l0 = hb.multi_launch_count()
    hb.compress_batch_host_multi(devs, nblk, pin_raw.array, r_off, r_len, pin_c.array, c_off, c_len, status, orders)
    assert hb.multi_launch_count() > l0
    assert (status == 0).all()
    for i, k in enumerate(idx):
        assert bytes(pin_c.array[int(c_off[i]): int(c_off[i]) + int(c_len[i])]) == want[k], (i, hex(orders1[k]))
    st = hb.multi_last_stats()
    assert [s["device"] for s in st] == devs
    if phased and len(devs) > 1:                                   # static partition: the C rule the call itself used
        ranges = shard.partition_blocks(r_len, len(devs))
        assert [(s["first_blk"], s["first_blk"] + s["nblk"]) for s in st] == ranges
    else:                                                          # chunks taken from the shared queue
        assert sum(s["nblk"] for s in st) == nblk
    assert sum(s["in_bytes"] for s in st) == int(r_len.astype(np.uint64).sum())
    assert sum(s["out_bytes"] for s in st) == int(c_len.astype(np.uint64).sum())
    # ---- decode them back over all devices
    pin_out = hb.PinnedArray(pin_raw.nbytes)
    out_len = r_len.copy()
    status[:] = 77
    hb.uncompress_batch_host_multi(devs, nblk, pin_c.array, c_off, c_len, pin_out.array, r_off, out_len, status)
    assert (status == 0).all() and (out_len == r_len).all()
    assert np.array_equal(pin_out.array, pin_raw.array)
    st = hb.multi_last_stats()
    assert all(s["wall_ms"] > 0 and s["kernel_ms"] > 0 and s["d2h_ms"] > 0 for s in st if s["nblk"])
    if phased and len(devs) > 1:
        assert all(s["d2h_phase_ms"] > 0 for s in st)
    hb.multi_set_phased(None)


def test_multi_bad_block_does_not_strand_the_other_devices(oracle):
    """A malformed stream fails alone (per-block status); more devices than blocks leaves some devices idle."""
    devs = _devices()
    raw = synth.qual_block(1, 5000).tobytes()
    good = oracle.compress(raw, 0)
    streams = [good, b"\x00\x05garbage!!", good]
    in_len = np.array([len(s) for s in streams], np.uint32)
    in_off = np.zeros(3, np.uint64); in_off[1:] = np.cumsum(in_len[:-1].astype(np.uint64))
    ib = np.frombuffer(b"".join(streams), np.uint8).copy()
    ob = np.zeros(3 * 5000, np.uint8)
    out_len = np.full(3, 5000, np.uint32)
    status = np.zeros(3, np.int32)
    hb.uncompress_batch_host_multi(devs, 3, ib, in_off, in_len, ob, np.arange(3, dtype=np.uint64) * 5000, out_len, status)
    assert status[0] == 0 and status[2] == 0 and status[1] != 0
    assert bytes(ob[:5000]) == raw and bytes(ob[10000:]) == raw
    # an empty batch and a duplicate device list
    hb.uncompress_batch_host_multi(devs, 0, ib, in_off, in_len, ob, in_off, out_len, status)
    with pytest.raises(RuntimeError):
        hb.uncompress_batch_host_multi([0, 0], 3, ib, in_off, in_len, ob, in_off, out_len, status)
