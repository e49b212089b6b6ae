"""ctypes loaders for the CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``  -- oracle/libhtsoracle.so, our C restatement (always buildable: plain gcc).
* ``RefLib``  -- oracle/_ref/libref.so, the unmodified reference C compiled by oracle/Makefile
                 (built in the authoring container; travels to the GPU box as a prebuilt file).

Nothing in htscodecs_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)

X_ORDER1, X_32, X_STRIPE, X_NOSZ, X_CAT, X_RLE, X_PACK = 0x01, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80


def _buf(b):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b) + (b"\0" if not len(b) else b""))


def mt_run(lib, streams, olens=None, orders=None, threads=1, reps=1, encode=False, method=0):
    """Time `reps` passes over `streams` on `threads` pthreads inside C (oracle/ref_mt.c).
    Returns (seconds, bytes produced)."""
    import numpy as np
    n = len(streams)
    ilen = np.array([len(s) for s in streams], np.uint32)
    off = np.zeros(n, np.uint64)
    off[1:] = np.cumsum(ilen[:-1].astype(np.uint64))
    base = np.frombuffer(b"".join(bytes(s) for s in streams) + b"\0" * 16, np.uint8)
    olen = np.array(olens if olens is not None else [0] * n, np.uint32)
    order = np.array(orders if orders is not None else [0] * n, np.int32)
    produced = C.c_uint64(0)
    lib.ref_mt_run.restype = C.c_double
    lib.ref_mt_run.argtypes = [C.c_void_p] * 5 + [C.c_int] * 5 + [C.POINTER(C.c_uint64)]
    t = lib.ref_mt_run(base.ctypes.data, off.ctypes.data, ilen.ctypes.data, olen.ctypes.data, order.ctypes.data,
                       n, threads, reps, 1 if encode else 0, method, C.byref(produced))
    return t, produced.value


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libhtsoracle.so")
    src = os.path.join(ORACLE_DIR, "hts_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libhtsoracle.so"], stdout=subprocess.DEVNULL)
    return so


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.ho_compress_bound.restype = C.c_uint
        L.ho_compress_bound.argtypes = [C.c_uint, C.c_int]
        L.ho_compress.argtypes = [u8p, C.c_uint32, u8p, u32p, C.c_int]
        L.ho_uncompress.argtypes = [u8p, C.c_uint32, u8p, u32p]
        L.ho_uncompress_4x8.argtypes = [u8p, C.c_uint32, u8p, u32p]
        L.ho_peek_size.argtypes = [u8p, C.c_uint32, u32p]
        L.ho_compress_bound_4x8.restype = C.c_uint
        L.ho_compress_bound_4x8.argtypes = [C.c_uint]
        L.ho_compress_4x8.argtypes = [u8p, C.c_uint32, u8p, u32p, C.c_int]
        L.ho_var_put_u32.argtypes = [u8p, C.c_uint32]
        L.ho_var_get_u32.argtypes = [u8p, u8p, u32p]

    def bound(self, n, order):
        return self.lib.ho_compress_bound(n, order)

    def compress(self, data, order):
        data = bytes(data)
        cap = self.bound(len(data), order) + 64
        out = (C.c_uint8 * cap)()
        osz = C.c_uint32(cap)
        rc = self.lib.ho_compress(_buf(data), len(data), out, C.byref(osz), order)
        if rc != 0:
            return None
        return C.string_at(out, osz.value)

    def peek_size(self, comp):
        v = C.c_uint32(0)
        rc = self.lib.ho_peek_size(_buf(comp), len(comp), C.byref(v))
        return v.value if rc == 0 else None

    def uncompress(self, comp, ulen=None):
        comp = bytes(comp)
        if ulen is None:
            ulen = self.peek_size(comp)
            if ulen is None:
                return None
        out = (C.c_uint8 * max(1, ulen))()
        osz = C.c_uint32(ulen)
        rc = self.lib.ho_uncompress(_buf(comp), len(comp), out, C.byref(osz))
        if rc != 0:
            return None
        return C.string_at(out, osz.value)

    def uncompress_4x8(self, comp):
        comp = bytes(comp)
        if len(comp) < 9:
            return None
        ulen = int.from_bytes(comp[5:9], "little")
        out = (C.c_uint8 * max(1, ulen))()
        osz = C.c_uint32(ulen)
        rc = self.lib.ho_uncompress_4x8(_buf(comp), len(comp), out, C.byref(osz))
        if rc != 0:
            return None
        return C.string_at(out, osz.value)

    def compress_4x8(self, data, order):
        data = bytes(data)
        cap = self.lib.ho_compress_bound_4x8(len(data)) + 64
        out = (C.c_uint8 * cap)()
        osz = C.c_uint32(cap)
        rc = self.lib.ho_compress_4x8(_buf(data), len(data), out, C.byref(osz), order)
        if rc != 0:
            return None
        return C.string_at(out, osz.value)

    def var_put(self, v):
        b = (C.c_uint8 * 8)()
        n = self.lib.ho_var_put_u32(b, v)
        return bytes(b[:n])


class RefLib:
    """The unmodified reference, for pinning the oracle and as the CPU baseline."""

    PATH = os.path.join(ORACLE_DIR, "_ref", "libref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = C.CDLL(self.PATH)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        L = self.lib
        L.rans_compress_bound_4x16.restype = C.c_uint
        L.rans_compress_bound_4x16.argtypes = [C.c_uint, C.c_int]
        L.rans_compress_to_4x16.restype = C.c_void_p
        L.rans_compress_to_4x16.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, u32p, C.c_int]
        L.rans_uncompress_to_4x16.restype = C.c_void_p
        L.rans_uncompress_to_4x16.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, u32p]
        L.rans_compress.restype = C.c_void_p
        L.rans_compress.argtypes = [C.c_void_p, C.c_uint, u32p, C.c_int]
        L.rans_uncompress.restype = C.c_void_p
        L.rans_uncompress.argtypes = [C.c_void_p, C.c_uint, u32p]

    def bound(self, n, order):
        return self.lib.rans_compress_bound_4x16(n, order)

    def compress(self, data, order):
        data = bytes(data)
        cap = self.bound(len(data), order)
        out = (C.c_uint8 * cap)()
        osz = C.c_uint32(cap)
        r = self.lib.rans_compress_to_4x16(_buf(data), len(data), out, C.byref(osz), order)
        if not r:
            return None
        return C.string_at(out, osz.value)

    def uncompress(self, comp, ulen):
        comp = bytes(comp)
        out = (C.c_uint8 * max(1, ulen))()
        osz = C.c_uint32(ulen)
        r = self.lib.rans_uncompress_to_4x16(_buf(comp), len(comp), out, C.byref(osz))
        if not r:
            return None
        return C.string_at(out, osz.value)

    def compress_4x8(self, data, order):
        data = bytes(data)
        osz = C.c_uint32(0)
        r = self.lib.rans_compress(_buf(data), len(data), C.byref(osz), order)
        if not r:
            return None
        res = C.string_at(r, osz.value)
        self.libc.free(r)
        return res

    def uncompress_4x8(self, comp):
        comp = bytes(comp)
        osz = C.c_uint32(0)
        r = self.lib.rans_uncompress(_buf(comp), len(comp), C.byref(osz))
        if not r:
            return None
        res = C.string_at(r, osz.value)
        self.libc.free(r)
        return res
