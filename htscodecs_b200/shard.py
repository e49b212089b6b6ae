"""Block sharding across the GPUs of one box.

CRAM blocks are independent, so the multi-GPU form of the hot path is a partition of the block
list: rank r (one process per GPU) codes the contiguous range ``ranges[r]`` on its own device and
the per-block results (sizes, status) are gathered on the host.  There is no exchange step on the
data path, hence no NCCL collective in it; the gather below moves 8 bytes per block and runs on
whatever process group the caller initialised (gloo on CPU, nccl on GPUs).

The codec call is injected (``worker``) so the same host logic is exercised by the CPU-only
world_size-2 gloo test (tests/test_shard.py, where the worker is a CPU checker) and by bench.py / the GPU tests
(worker = Context.uncompress_batch_host / compress_batch_host).
"""
import numpy as np


def partition_blocks(weights, world):
    """Split blocks [0, n) into `world` contiguous ranges balanced on `weights` (uncompressed
    bytes per block).  Returns a list of (lo, hi); ranges may be empty when n < world.

    This is the library's own rule (hts_b200_partition, the one the multi-device C calls use): range r
    ends at the block boundary nearest to (r + 1) / world of the grand total, which keeps every range
    within one block of the ideal.  Pure host arithmetic (no GPU needed)."""
    import ctypes as C
    from . import load_library
    if world <= 0:
        raise ValueError("world must be positive")
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.uint32))
    n = int(w.size)
    cuts = (C.c_int * (world + 1))()
    rc = load_library().hts_b200_partition(n, w.ctypes.data if n else None, world, cuts)
    if rc != 0:
        raise ValueError("hts_b200_partition failed")
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_results(out_len_local, status_local, ranges, rank, world, dist=None, device=None):
    """Host-side gather of the per-block (out_len, status) arrays of every rank.

    `dist` is torch.distributed (initialised) or None for world == 1.  Returns full-length
    (out_len uint32, status int32) arrays on every rank."""
    n = ranges[-1][1]
    out_len = np.zeros(n, np.uint32)
    status = np.zeros(n, np.int32)
    lo, hi = ranges[rank]
    out_len[lo:hi] = out_len_local
    status[lo:hi] = status_local
    if world == 1 or dist is None:
        return out_len, status
    import torch
    # one fixed-size slot per rank (ranges differ by at most a few blocks)
    slot = max(1, max(h - l for l, h in ranges))
    mine = torch.zeros(2 * slot, dtype=torch.int64)
    mine[: hi - lo] = torch.from_numpy(out_len_local.astype(np.int64))
    mine[slot: slot + hi - lo] = torch.from_numpy(status_local.astype(np.int64))
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    for r, (l, h) in enumerate(ranges):
        p = parts[r].cpu().numpy()
        out_len[l:h] = p[: h - l].astype(np.uint32)
        status[l:h] = p[slot: slot + h - l].astype(np.int32)
    return out_len, status


def exclusive_offsets(sizes, align=1):
    """Offsets of a packed arena holding blocks of `sizes` bytes (the encode-side gather: every
    rank learns where each compressed block would sit in the concatenated output)."""
    s = (np.asarray(sizes, dtype=np.uint64) + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    off = np.zeros(s.size, np.uint64)
    if s.size > 1:
        off[1:] = np.cumsum(s[:-1])
    return off


def run_sharded(worker, weights, rank, world, dist=None, device=None):
    """Partition, run `worker(lo, hi) -> (out_len[hi-lo], status[hi-lo])` on this rank's range and
    gather.  Returns (ranges, out_len, status) with full-length arrays."""
    ranges = partition_blocks(weights, world)
    lo, hi = ranges[rank]
    if hi > lo:
        ol, st = worker(lo, hi)
    else:
        ol, st = np.zeros(0, np.uint32), np.zeros(0, np.int32)
    ol = np.asarray(ol, np.uint32)
    st = np.asarray(st, np.int32)
    assert ol.size == hi - lo and st.size == hi - lo
    out_len, status = gather_results(ol, st, ranges, rank, world, dist, device)
    return ranges, out_len, status
