"""htscodecs_b200 -- Python-side mirror of the C ABI in include/htscodecs_b200.h.

The product is ``libhtscodecs_b200.so`` (hand-written sm_100a CUDA kernels behind the reference's
own C entry points plus batched ones).  This module only binds it with ctypes so that tests and
bench.py can call it the way a C caller would; no codec work happens in Python and there is no
CPU fallback: if the library is missing, or no sm_100 GPU is usable, calls raise / return None.

Function names follow the reference (htscodecs/rANS_static4x16.h:40-50, rANS_static.h:40-43).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhtscodecs_b200.so")

RANS_ORDER_1, RANS_ORDER_X32, RANS_ORDER_STRIPE, RANS_ORDER_NOSZ = 0x01, 0x04, 0x08, 0x10
RANS_ORDER_CAT, RANS_ORDER_RLE, RANS_ORDER_PACK = 0x20, 0x40, 0x80
RANS4x16, RANS4x8 = 0, 1
ORDER_RANS4x8 = 0x40000000          # OR into a batched encoder's order[i]: legacy rANS 4x8 codec

EXPORTS = [
    "rans_compress_bound_4x16", "rans_compress_to_4x16", "rans_compress_4x16",
    "rans_uncompress_to_4x16", "rans_uncompress_4x16", "rans_uncompress", "rans_compress",
    "hts_b200_compress_bound_4x8",
    "hts_b200_create", "hts_b200_destroy", "hts_b200_last_error", "hts_b200_launch_count",
    "hts_b200_stream", "hts_b200_scratch_bytes", "hts_b200_uncompress_batch_dev", "hts_b200_uncompress_batch_host",
    "hts_b200_compress_batch_dev", "hts_b200_compress_batch_dev_async", "hts_b200_compress_batch_host", "rans4x16_uncompress_batch",
    "rans4x16_compress_batch", "rans4x16_compress_best_batch", "hts_b200_peek_size", "hts_b200_host_alloc", "hts_b200_host_free",
    "hts_b200_set_copy_duplex", "hts_b200_plan_chunks",
    "hts_b200_uncompress_batch_host_multi", "hts_b200_compress_batch_host_multi", "hts_b200_multi_set_phased",
    "hts_b200_multi_last_stats", "hts_b200_multi_launch_count", "hts_b200_multi_last_error", "hts_b200_partition",
    "hts_b200_frames_scan", "hts_b200_frames_write",
]

_lib = None
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def load_library():
    """dlopen the in-tree library (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(htscodecs_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int32, C.c_uint32
    lib.rans_compress_bound_4x16.restype = C.c_uint
    lib.rans_compress_bound_4x16.argtypes = [C.c_uint, C.c_int]
    lib.rans_compress_to_4x16.restype = vp
    lib.rans_compress_to_4x16.argtypes = [vp, C.c_uint, vp, _u32p, C.c_int]
    lib.rans_compress_4x16.restype = vp
    lib.rans_compress_4x16.argtypes = [vp, C.c_uint, _u32p, C.c_int]
    lib.rans_uncompress_to_4x16.restype = vp
    lib.rans_uncompress_to_4x16.argtypes = [vp, C.c_uint, vp, _u32p]
    lib.rans_uncompress_4x16.restype = vp
    lib.rans_uncompress_4x16.argtypes = [vp, C.c_uint, _u32p]
    lib.rans_uncompress.restype = vp
    lib.rans_uncompress.argtypes = [vp, C.c_uint, _u32p]
    lib.rans_compress.restype = vp
    lib.rans_compress.argtypes = [vp, C.c_uint, _u32p, C.c_int]
    lib.hts_b200_compress_bound_4x8.restype = C.c_uint
    lib.hts_b200_compress_bound_4x8.argtypes = [C.c_uint]
    lib.hts_b200_create.restype = vp
    lib.hts_b200_create.argtypes = [C.c_int]
    lib.hts_b200_destroy.argtypes = [vp]
    lib.hts_b200_last_error.restype = C.c_char_p
    lib.hts_b200_last_error.argtypes = [vp]
    lib.hts_b200_launch_count.restype = C.c_ulonglong
    lib.hts_b200_launch_count.argtypes = [vp]
    lib.hts_b200_stream.restype = vp
    lib.hts_b200_stream.argtypes = [vp]
    lib.hts_b200_scratch_bytes.restype = C.c_size_t
    lib.hts_b200_scratch_bytes.argtypes = [vp]
    batch = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.hts_b200_uncompress_batch_dev.argtypes = batch + [C.c_int]
    lib.hts_b200_uncompress_batch_host.argtypes = batch
    lib.hts_b200_compress_batch_dev.argtypes = batch + [C.c_int]
    lib.hts_b200_compress_batch_host.argtypes = batch
    lib.hts_b200_compress_batch_dev_async.argtypes = batch + [vp, vp]
    lib.rans4x16_uncompress_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    lib.rans4x16_compress_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.rans4x16_compress_best_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp]
    lib.hts_b200_peek_size.argtypes = [vp, u32, C.c_int, _u32p]
    lib.hts_b200_host_alloc.restype = vp
    lib.hts_b200_host_alloc.argtypes = [C.c_size_t]
    lib.hts_b200_host_free.argtypes = [vp]
    lib.hts_b200_set_copy_duplex.argtypes = [vp, C.c_int]
    lib.hts_b200_plan_chunks.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    multi = [C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.hts_b200_uncompress_batch_host_multi.argtypes = multi
    lib.hts_b200_compress_batch_host_multi.argtypes = multi
    lib.hts_b200_multi_set_phased.argtypes = [C.c_int]
    lib.hts_b200_multi_last_stats.argtypes = [vp, C.c_int]
    lib.hts_b200_multi_launch_count.restype = C.c_ulonglong
    lib.hts_b200_multi_last_error.restype = C.c_char_p
    lib.hts_b200_partition.argtypes = [C.c_int, vp, C.c_int, vp]
    lib.hts_b200_frames_scan.restype = C.c_long
    lib.hts_b200_frames_scan.argtypes = [vp, C.c_size_t, C.c_int, C.c_long, vp, vp, vp, vp, vp, u32]
    lib.hts_b200_frames_write.restype = C.c_size_t
    lib.hts_b200_frames_write.argtypes = [vp, C.c_size_t, C.c_long, vp, vp, vp, vp]
    _libc = C.CDLL(None)
    _libc.free.argtypes = [vp]
    lib._free = _libc.free
    _lib = lib
    return lib


# ------------------------------------------------------------------------------------------------
# drop-in single-block calls (what htslib would call)
# ------------------------------------------------------------------------------------------------
def _inbuf(data):
    data = bytes(data)
    return (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if data else b"\0"), len(data)


def rans_compress_bound_4x16(size, order):
    return load_library().rans_compress_bound_4x16(size, order)


def rans_compress_4x16(data, order):
    """bytes -> compressed bytes, or None on failure (the C call returns NULL)."""
    lib = load_library()
    buf, n = _inbuf(data)
    osz = C.c_uint32(0)
    p = lib.rans_compress_4x16(buf, n, C.byref(osz), order)
    if not p:
        return None
    out = C.string_at(p, osz.value)
    lib._free(p)
    return out


def rans_compress_to_4x16(data, order, capacity=None):
    lib = load_library()
    buf, n = _inbuf(data)
    cap = lib.rans_compress_bound_4x16(n, order) if capacity is None else capacity
    out = (C.c_uint8 * max(1, cap))()
    osz = C.c_uint32(cap)
    p = lib.rans_compress_to_4x16(buf, n, out, C.byref(osz), order)
    return bytes(out[: osz.value]) if p else None


def rans_uncompress_4x16(data):
    lib = load_library()
    buf, n = _inbuf(data)
    osz = C.c_uint32(0)
    p = lib.rans_uncompress_4x16(buf, n, C.byref(osz))
    if not p:
        return None
    out = C.string_at(p, osz.value)
    lib._free(p)
    return out


def rans_uncompress_to_4x16(data, out_size):
    """Decode into a caller-sized buffer (required for X_NOSZ streams)."""
    lib = load_library()
    buf, n = _inbuf(data)
    out = (C.c_uint8 * max(1, out_size))()
    osz = C.c_uint32(out_size)
    p = lib.rans_uncompress_to_4x16(buf, n, out, C.byref(osz))
    return bytes(out[: osz.value]) if p else None


def rans_uncompress(data):
    """Legacy rANS 4x8 (CRAM 3.0) decode."""
    lib = load_library()
    buf, n = _inbuf(data)
    osz = C.c_uint32(0)
    p = lib.rans_uncompress(buf, n, C.byref(osz))
    if not p:
        return None
    out = C.string_at(p, osz.value)
    lib._free(p)
    return out


def rans_compress(data, order):
    """Legacy rANS 4x8 (CRAM 3.0) encode; None on failure (empty input included)."""
    lib = load_library()
    buf, n = _inbuf(data)
    osz = C.c_uint32(0)
    p = lib.rans_compress(buf, n, C.byref(osz), order)
    if not p:
        return None
    out = C.string_at(p, osz.value)
    lib._free(p)
    return out


def peek_size(data, method=RANS4x16):
    lib = load_library()
    buf, n = _inbuf(data)
    v = C.c_uint32(0)
    return v.value if lib.hts_b200_peek_size(buf, n, method, C.byref(v)) == 0 else None


# ------------------------------------------------------------------------------------------------
# batched calls
# ------------------------------------------------------------------------------------------------
def _ptr(a):
    """Address of a numpy array, a torch tensor (host or device), an int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


class PinnedArray:
    """A numpy view over cudaHostAlloc'ed memory (freed with the object)."""

    def __init__(self, nbytes):
        self.lib = load_library()
        self.nbytes = int(nbytes)
        self.ptr = self.lib.hts_b200_host_alloc(max(1, self.nbytes))
        if not self.ptr:
            raise MemoryError("cudaHostAlloc failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr))[: self.nbytes]

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.hts_b200_host_free(self.ptr)
            self.ptr = None


class DevStats(C.Structure):
    """hts_b200_dev_stats (include/htscodecs_b200.h)."""
    _fields_ = [("device", C.c_int), ("first_blk", C.c_int), ("nblk", C.c_int), ("pad", C.c_int),
                ("in_bytes", C.c_uint64), ("out_bytes", C.c_uint64),
                ("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_phase_ms", C.c_double), ("wait_ms", C.c_double), ("d2h_phase_ms", C.c_double),
                ("wall_ms", C.c_double)]

    def as_dict(self):
        return {k: (round(getattr(self, k), 3) if isinstance(getattr(self, k), float) else getattr(self, k))
                for k, _ in self._fields_ if k != "pad"}


def _multi(fn, devices, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, last):
    lib = load_library()
    dev = np.ascontiguousarray(np.asarray(devices, np.int32))
    rc = fn(len(dev), dev.ctypes.data, nblk, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off),
            _ptr(out_len), _ptr(status), _ptr(last))
    if rc != 0:
        raise RuntimeError("htscodecs_b200 multi-device call failed: " + lib.hts_b200_multi_last_error().decode())


def uncompress_batch_host_multi(devices, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, method=None):
    """hts_b200_uncompress_batch_host_multi: one host-buffer batch over several GPUs of this process."""
    _multi(load_library().hts_b200_uncompress_batch_host_multi, devices, nblk, in_base, in_off, in_len, out_base,
           out_off, out_len, status, method)


def compress_batch_host_multi(devices, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, order):
    _multi(load_library().hts_b200_compress_batch_host_multi, devices, nblk, in_base, in_off, in_len, out_base,
           out_off, out_len, status, order)


def multi_set_phased(phased):
    """True: phased copies across devices, False: full duplex per device, None: let the library measure (default)."""
    load_library().hts_b200_multi_set_phased(-1 if phased is None else (1 if phased else 0))


def multi_last_stats():
    """Per-device breakdown of the last multi-device call: list of dicts."""
    lib = load_library()
    arr = (DevStats * 64)()
    n = lib.hts_b200_multi_last_stats(C.addressof(arr), 64)
    return [arr[i].as_dict() for i in range(min(n, 64))]


def multi_launch_count():
    return int(load_library().hts_b200_multi_launch_count())


def frames_scan(buf, method=RANS4x16, out_align=1):
    """`[u32 clen][stream]...` (the reference test programs' file format) -> (in_off, in_len, out_off,
    out_len, out_total) numpy arrays for the host-buffer decode calls; None when the buffer is malformed."""
    lib = load_library()
    b = np.frombuffer(bytes(buf), np.uint8) if not isinstance(buf, np.ndarray) else buf
    addr = b.ctypes.data if b.size else None
    n = lib.hts_b200_frames_scan(addr, b.size, method, 0, None, None, None, None, None, out_align)
    if n < 0:
        return None
    in_off, out_off = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    in_len, out_len = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    total = C.c_uint64(0)
    lib.hts_b200_frames_scan(addr, b.size, method, n, _ptr(in_off), _ptr(in_len), _ptr(out_off), _ptr(out_len),
                             C.addressof(total), out_align)
    return in_off, in_len, out_off, out_len, int(total.value)


def frames_write(src, src_off, src_len, status=None):
    """Encoded streams -> `[u32 clen][stream]...` bytes (failed blocks skipped)."""
    lib = load_library()
    n = len(src_len)
    src_off = np.ascontiguousarray(src_off, np.uint64)
    src_len = np.ascontiguousarray(src_len, np.uint32)
    st = None if status is None else np.ascontiguousarray(status, np.int32)
    need = lib.hts_b200_frames_write(None, 0, n, _ptr(src), _ptr(src_off), _ptr(src_len), _ptr(st))
    dst = np.zeros(max(1, need), np.uint8)
    got = lib.hts_b200_frames_write(_ptr(dst), need, n, _ptr(src), _ptr(src_off), _ptr(src_len), _ptr(st))
    assert got == need
    return dst[:need].tobytes()


class Context:
    """hts_b200_ctx: one per host thread and device."""

    def __init__(self, device=-1):
        self.lib = load_library()
        self.h = self.lib.hts_b200_create(device)
        if not self.h:
            raise RuntimeError("hts_b200_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")

    def close(self):
        if self.h:
            self.lib.hts_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.hts_b200_launch_count(self.h))

    @property
    def stream(self):
        return self.lib.hts_b200_stream(self.h)

    @property
    def scratch_bytes(self):
        """Device memory held for work lists, scratch arenas and staging."""
        return int(self.lib.hts_b200_scratch_bytes(self.h))

    def set_copy_duplex(self, full):
        """False: send every input of a host-buffer call before fetching any result (half duplex)."""
        self.lib.hts_b200_set_copy_duplex(self.h, 1 if full else 0)

    def last_error(self):
        return self.lib.hts_b200_last_error(self.h).decode()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("htscodecs_b200 batch call failed: " + self.last_error())

    # -- device-resident (all arguments are device pointers / CUDA tensors) ----------------------
    @staticmethod
    def _wait_producers():
        """The context's stream is non-blocking, so it does not wait for work torch queued on ITS current
        stream (torch.full / .cuda() / fill_ of the arguments).  Drain that stream before the library reads them;
        on an idle stream this is a few microseconds."""
        import sys
        torch = sys.modules.get("torch")
        if torch is not None and torch.cuda.is_initialized():
            torch.cuda.current_stream().synchronize()

    def uncompress_batch_dev(self, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status,
                             method=None, sync=True):
        self._wait_producers()
        self._check(self.lib.hts_b200_uncompress_batch_dev(
            self.h, nblk, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off),
            _ptr(out_len), _ptr(status), _ptr(method), 1 if sync else 0))

    def compress_batch_dev(self, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, order,
                           sync=True):
        self._wait_producers()
        self._check(self.lib.hts_b200_compress_batch_dev(
            self.h, nblk, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off),
            _ptr(out_len), _ptr(status), _ptr(order), 1 if sync else 0))

    def compress_batch_dev_async(self, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, order,
                                 host_in_len, host_order):
        """Device-resident encode that never synchronises the context's stream: host_in_len / host_order are numpy
        copies of in_len / order."""
        self._wait_producers()
        self._check(self.lib.hts_b200_compress_batch_dev_async(
            self.h, nblk, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off),
            _ptr(out_len), _ptr(status), _ptr(order), _ptr(host_in_len), _ptr(host_order)))

    # -- host-resident (numpy arrays; pinned or pageable) ------------------------------------------
    def uncompress_batch_host(self, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status,
                              method=None):
        self._check(self.lib.hts_b200_uncompress_batch_host(
            self.h, nblk, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off),
            _ptr(out_len), _ptr(status), _ptr(method)))

    def compress_batch_host(self, nblk, in_base, in_off, in_len, out_base, out_off, out_len, status, order):
        self._check(self.lib.hts_b200_compress_batch_host(
            self.h, nblk, _ptr(in_base), _ptr(in_off), _ptr(in_len), _ptr(out_base), _ptr(out_off),
            _ptr(out_len), _ptr(status), _ptr(order)))

    # -- convenience: lists of bytes in, lists of bytes out ----------------------------------------
    def uncompress_many(self, streams, sizes=None, methods=None):
        """Decode a list of streams.  sizes[i] = expected size (needed for X_NOSZ); returns
        (list of bytes-or-None, status array)."""
        n = len(streams)
        if n == 0:
            return [], np.zeros(0, np.int32)
        if sizes is None:
            sizes = [peek_size(s, methods[i] if methods is not None else RANS4x16) for i, s in enumerate(streams)]
        caps = np.array([0 if s is None else s for s in sizes], np.uint32)
        in_len = np.array([len(s) for s in streams], np.uint32)
        in_off = np.zeros(n, np.uint64)
        out_off = np.zeros(n, np.uint64)
        if n > 1:
            in_off[1:] = np.cumsum((in_len[:-1].astype(np.uint64) + 15) // 16 * 16)
            out_off[1:] = np.cumsum((caps[:-1].astype(np.uint64) + 15) // 16 * 16)
        ib = np.zeros(int(in_off[-1] + in_len[-1]) + 16, np.uint8)
        for i, s in enumerate(streams):
            ib[int(in_off[i]): int(in_off[i]) + len(s)] = np.frombuffer(bytes(s), np.uint8)
        ob = np.zeros(int(out_off[-1] + caps[-1]) + 16, np.uint8)
        out_len = caps.copy()
        status = np.zeros(n, np.int32)
        meth = None if methods is None else np.array(methods, np.uint8)
        self.uncompress_batch_host(n, ib, in_off, in_len, ob, out_off, out_len, status, meth)
        res = [bytes(ob[int(out_off[i]): int(out_off[i]) + int(out_len[i])]) if status[i] == 0 else None
               for i in range(n)]
        return res, status

    def uncompress_many_dev(self, streams, sizes, methods=None):
        """As uncompress_many, through the device-resident entry point (one call, whatever the batch
        size: large batches reach the high-occupancy kernel variants).  Needs torch."""
        import torch
        n = len(streams)
        caps = np.array(sizes, np.uint32)
        in_len = np.array([len(s) for s in streams], np.uint32)
        in_off = np.zeros(n, np.uint64)
        out_off = np.zeros(n, np.uint64)
        if n > 1:
            in_off[1:] = np.cumsum((in_len[:-1].astype(np.uint64) + 15) // 16 * 16)
            out_off[1:] = np.cumsum((caps[:-1].astype(np.uint64) + 15) // 16 * 16)
        ib = np.zeros(int(in_off[-1] + in_len[-1]) + 16, np.uint8)
        for i, s in enumerate(streams):
            ib[int(in_off[i]): int(in_off[i]) + len(s)] = np.frombuffer(bytes(s), np.uint8)
        d_in = torch.from_numpy(ib).cuda()
        d_out = torch.zeros(int(out_off[-1] + caps[-1]) + 16, dtype=torch.uint8, device="cuda")
        d_len = torch.from_numpy(caps.view(np.int32).copy()).cuda()
        d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_m = None if methods is None else torch.from_numpy(np.array(methods, np.uint8)).cuda()
        self.uncompress_batch_dev(n, d_in, torch.from_numpy(in_off.view(np.int64)).cuda(),
                                  torch.from_numpy(in_len.view(np.int32)).cuda(), d_out,
                                  torch.from_numpy(out_off.view(np.int64)).cuda(), d_len, d_st, d_m)
        ob, out_len, status = d_out.cpu().numpy(), d_len.cpu().numpy().view(np.uint32), d_st.cpu().numpy()
        res = [bytes(ob[int(out_off[i]): int(out_off[i]) + int(min(out_len[i], caps[i]))]) if status[i] == 0 else None
               for i in range(n)]
        return res, status

    def compress_best_many(self, blocks, methods):
        """Try-all-methods encode (the selection loop of tokenise_name3.c:1246-1299): every block is
        coded with every order in `methods` in one batched pass and the first strictly smallest
        stream is kept.  Returns (streams, best_orders, status)."""
        n = len(blocks)
        if n == 0:
            return [], np.zeros(0, np.int32), np.zeros(0, np.int32)
        lib = self.lib
        bufs = [np.frombuffer(bytes(b), np.uint8) if len(b) else np.zeros(1, np.uint8) for b in blocks]
        in_size = np.array([len(b) for b in blocks], np.uint32)
        meth = np.array(methods, np.int32)
        caps = np.array([max(lib.rans_compress_bound_4x16(int(in_size[i]), int(m)) for m in meth) for i in range(n)], np.uint32)
        outs = [np.zeros(int(c), np.uint8) for c in caps]
        in_ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        out_ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        out_size = caps.copy()
        best = np.full(n, -1, np.int32)
        status = np.zeros(n, np.int32)
        self._check(lib.rans4x16_compress_best_batch(
            self.h, n, C.cast(in_ptrs, C.c_void_p), _ptr(in_size), C.cast(out_ptrs, C.c_void_p), _ptr(out_size),
            _ptr(meth), len(meth), _ptr(best), _ptr(status)))
        res = [bytes(outs[i][: int(out_size[i])]) if status[i] == 0 else None for i in range(n)]
        return res, best, status

    def compress_many(self, blocks, orders):
        n = len(blocks)
        if n == 0:
            return [], np.zeros(0, np.int32)
        lib = self.lib
        in_len = np.array([len(b) for b in blocks], np.uint32)
        order = np.array(orders, np.int32)
        caps = np.array([lib.hts_b200_compress_bound_4x8(int(in_len[i])) if int(order[i]) & ORDER_RANS4x8
                         else lib.rans_compress_bound_4x16(int(in_len[i]), int(order[i])) for i in range(n)], np.uint32)
        in_off = np.zeros(n, np.uint64)
        out_off = np.zeros(n, np.uint64)
        if n > 1:
            in_off[1:] = np.cumsum((in_len[:-1].astype(np.uint64) + 15) // 16 * 16)
            out_off[1:] = np.cumsum((caps[:-1].astype(np.uint64) + 15) // 16 * 16)
        ib = np.zeros(int(in_off[-1] + in_len[-1]) + 16, np.uint8)
        for i, b in enumerate(blocks):
            ib[int(in_off[i]): int(in_off[i]) + len(b)] = np.frombuffer(bytes(b), np.uint8)
        ob = np.zeros(int(out_off[-1] + caps[-1]) + 16, np.uint8)
        out_len = caps.copy()
        status = np.zeros(n, np.int32)
        self.compress_batch_host(n, ib, in_off, in_len, ob, out_off, out_len, status, order)
        res = [bytes(ob[int(out_off[i]): int(out_off[i]) + int(out_len[i])]) if status[i] == 0 else None
               for i in range(n)]
        return res, status
