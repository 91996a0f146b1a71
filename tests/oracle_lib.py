"""ctypes bindings for the CPU checkers (test infrastructure only).

* ``Oracle``  -> oracle/liboracle.so, our C restatement (oracle/bra_oracle.c).
* ``RefLib``  -> oracle/_ref/libbra_ref.so, the reference's own encoder/CRC sources
  compiled in place by oracle/Makefile (present when it was built in the dev
  container; it travels to the GPU box as a prebuilt file).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libbra_ref.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


def _buf(b):
    """bytes/bytearray/np.uint8 array -> (ctypes pointer, keepalive, length)."""
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b, dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)[:0]
    return a.ctypes.data_as(u8p), a, int(a.size)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "bra_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class Oracle:
    def __init__(self):
        self.lib = L = C.CDLL(build_oracle())
        L.ora_crc32c.restype = C.c_uint32
        L.ora_crc32c.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.ora_crc32c_combine.restype = C.c_uint32
        L.ora_crc32c_combine.argtypes = [C.c_uint32] * 3
        for f in (L.ora_bwt_encode, L.ora_bwt_encode_naive):
            f.restype = C.c_int
            f.argtypes = [u8p, C.c_uint32, u32p, u8p]
        L.ora_bwt_decode.restype = C.c_int
        L.ora_bwt_decode.argtypes = [u8p, C.c_uint32, C.c_uint32, u8p]
        for f in (L.ora_mtf_encode, L.ora_mtf_decode):
            f.restype = None
            f.argtypes = [u8p, C.c_size_t, u8p]
        L.ora_rle_encode_size.restype = C.c_size_t
        L.ora_rle_encode_size.argtypes = [u8p, C.c_size_t]
        L.ora_rle_decode_size.restype = C.c_size_t
        L.ora_rle_decode_size.argtypes = [u8p, C.c_size_t]
        for f in (L.ora_rle_encode, L.ora_rle_decode):
            f.restype = C.c_int
            f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ora_huffman_lengths.restype = C.c_int
        L.ora_huffman_lengths.argtypes = [u32p, u8p]
        L.ora_huffman_canonical.restype = None
        L.ora_huffman_canonical.argtypes = [u8p, u32p]
        L.ora_huffman_encode.restype = C.c_int
        L.ora_huffman_encode.argtypes = [u8p, C.c_uint32, u8p, u8p, C.c_size_t, u32p]
        L.ora_huffman_decode.restype = C.c_int
        L.ora_huffman_decode.argtypes = [u8p, u8p, C.c_uint32, C.c_uint32, u8p]
        L.ora_encode_block.restype = C.c_int
        L.ora_encode_block.argtypes = [u8p, C.c_uint32, u8p, u8p, C.c_size_t, u32p, C.c_int]
        L.ora_decode_block.restype = C.c_int
        L.ora_decode_block.argtypes = [u8p, u8p, u8p, C.c_size_t, u32p]

    # -- stage wrappers: bytes in, bytes out -------------------------------------------
    def crc32c(self, data, prev=0):
        p, keep, n = _buf(data)
        return int(self.lib.ora_crc32c(C.cast(p, C.c_void_p), n, prev))

    def crc32c_combine(self, a, b, len_b):
        return int(self.lib.ora_crc32c_combine(a, b, len_b & 0xFFFFFFFF))

    def bwt_encode(self, data, naive=False):
        p, keep, n = _buf(data)
        out = np.empty(n, dtype=np.uint8)
        pi = C.c_uint32(0)
        f = self.lib.ora_bwt_encode_naive if naive else self.lib.ora_bwt_encode
        if f(p, n, C.byref(pi), out.ctypes.data_as(u8p)) != 0:
            return None, None
        return out.tobytes(), int(pi.value)

    def bwt_decode(self, data, primary):
        p, keep, n = _buf(data)
        out = np.empty(n, dtype=np.uint8)
        if self.lib.ora_bwt_decode(p, n, primary, out.ctypes.data_as(u8p)) != 0:
            return None
        return out.tobytes()

    def mtf_encode(self, data):
        p, keep, n = _buf(data)
        out = np.empty(n, dtype=np.uint8)
        self.lib.ora_mtf_encode(p, n, out.ctypes.data_as(u8p))
        return out.tobytes()

    def mtf_decode(self, data):
        p, keep, n = _buf(data)
        out = np.empty(n, dtype=np.uint8)
        self.lib.ora_mtf_decode(p, n, out.ctypes.data_as(u8p))
        return out.tobytes()

    def rle_encode(self, data):
        p, keep, n = _buf(data)
        cap = n + n // 128 + 2
        out = np.empty(cap, dtype=np.uint8)
        on = C.c_size_t(0)
        if self.lib.ora_rle_encode(p, n, out.ctypes.data_as(u8p), cap, C.byref(on)) != 0:
            return None
        return out[: on.value].tobytes()

    def rle_decode_size(self, data):
        p, keep, n = _buf(data)
        return int(self.lib.ora_rle_decode_size(p, n))

    def rle_decode(self, data):
        p, keep, n = _buf(data)
        s = int(self.lib.ora_rle_decode_size(p, n))
        if s == 0:
            return None
        out = np.empty(s, dtype=np.uint8)
        on = C.c_size_t(0)
        if self.lib.ora_rle_decode(p, n, out.ctypes.data_as(u8p), s, C.byref(on)) != 0:
            return None
        return out[: on.value].tobytes()

    def huffman_lengths(self, freq):
        f = np.ascontiguousarray(freq, dtype=np.uint32)
        out = np.zeros(256, dtype=np.uint8)
        if self.lib.ora_huffman_lengths(f.ctypes.data_as(u32p), out.ctypes.data_as(u8p)) != 0:
            return None
        return out

    def huffman_canonical(self, lengths):
        l = np.ascontiguousarray(lengths, dtype=np.uint8)
        out = np.zeros(256, dtype=np.uint32)
        self.lib.ora_huffman_canonical(l.ctypes.data_as(u8p), out.ctypes.data_as(u32p))
        return out

    def huffman_encode(self, data):
        """-> (lengths[256] bytes, payload bytes) or None when n == 0."""
        p, keep, n = _buf(data)
        cap = n * 32 + 16
        lengths = np.zeros(256, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.uint8)
        es = C.c_uint32(0)
        if self.lib.ora_huffman_encode(p, n, lengths.ctypes.data_as(u8p), out.ctypes.data_as(u8p), cap, C.byref(es)) != 0:
            return None
        return lengths.tobytes(), out[: es.value].tobytes()

    def huffman_decode(self, lengths, payload, orig_size, encoded_size=None):
        l = np.frombuffer(bytes(lengths), dtype=np.uint8).copy()
        p, keep, n = _buf(payload)
        if encoded_size is None:
            encoded_size = n
        out = np.zeros(max(orig_size, 1), dtype=np.uint8)
        if self.lib.ora_huffman_decode(l.ctypes.data_as(u8p), p, encoded_size, orig_size, out.ctypes.data_as(u8p)) != 0:
            return None
        return out[:orig_size].tobytes()

    def encode_block(self, data, naive_bwt=False):
        """-> (hdr268 bytes, payload bytes, crc_raw)."""
        p, keep, n = _buf(data)
        cap = n + n // 64 + 1024
        hdr = np.zeros(268, dtype=np.uint8)
        out = np.zeros(cap, dtype=np.uint8)
        crc = C.c_uint32(0)
        if self.lib.ora_encode_block(p, n, hdr.ctypes.data_as(u8p), out.ctypes.data_as(u8p), cap, C.byref(crc), int(naive_bwt)) != 0:
            return None
        c = int.from_bytes(hdr[264:268].tobytes(), "little")
        return hdr.tobytes(), out[:c].tobytes(), int(crc.value)

    def decode_block(self, hdr, payload, cap):
        h = np.frombuffer(bytes(hdr), dtype=np.uint8).copy()
        p, keep, n = _buf(payload)
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        on = C.c_uint32(0)
        if self.lib.ora_decode_block(h.ctypes.data_as(u8p), p, out.ctypes.data_as(u8p), cap, C.byref(on)) != 0:
            return None
        return out[: on.value].tobytes()


class _HuffMeta(C.Structure):
    _pack_ = 1
    _fields_ = [("lengths", C.c_uint8 * 256), ("orig_size", C.c_uint32), ("encoded_size", C.c_uint32)]


class _HuffChunk(C.Structure):
    _fields_ = [("meta", _HuffMeta), ("data", u8p)]


class RefApi:
    """Binds the reference's public C API (reference src/encoders/*.h, src/utils/lib_bra_crc32c.h)
    on ANY shared library that exports it: oracle/_ref/libbra_ref.so (the reference itself) or
    the B200 drop-in library. Same wrapper for both, so parity tests read identically."""

    def __init__(self, path):
        self.path = path
        self.lib = L = C.CDLL(path)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        L.bra_crc32c.restype = C.c_uint32
        L.bra_crc32c.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        for name in ("bra_crc32c_table", "bra_crc32c_sse42"):
            f = getattr(L, name)
            f.restype = C.c_uint32
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.bra_crc32c_combine.restype = C.c_uint32
        L.bra_crc32c_combine.argtypes = [C.c_uint32] * 3
        L.bra_bwt_encode.restype = C.c_void_p
        L.bra_bwt_encode.argtypes = [u8p, C.c_uint32, u32p]
        L.bra_bwt_encode2.restype = C.c_bool
        L.bra_bwt_encode2.argtypes = [u8p, C.c_uint32, u32p, u8p]
        L.bra_bwt_decode.restype = C.c_void_p
        L.bra_bwt_decode.argtypes = [u8p, C.c_uint32, C.c_uint32]
        L.bra_bwt_decode2.restype = None
        L.bra_bwt_decode2.argtypes = [u8p, C.c_uint32, C.c_uint32, u32p, u8p]
        L.bra_mtf_encode.restype = C.c_void_p
        L.bra_mtf_encode.argtypes = [u8p, C.c_size_t]
        L.bra_mtf_encode2.restype = C.c_bool
        L.bra_mtf_encode2.argtypes = [u8p, C.c_size_t, u8p]
        L.bra_mtf_decode.restype = C.c_void_p
        L.bra_mtf_decode.argtypes = [u8p, C.c_size_t]
        L.bra_mtf_decode2.restype = None
        L.bra_mtf_decode2.argtypes = [u8p, C.c_size_t, u8p]
        L.bra_rle_encode.restype = C.c_bool
        L.bra_rle_encode.argtypes = [u8p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.bra_rle_decode.restype = C.c_bool
        L.bra_rle_decode.argtypes = [u8p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.bra_rle_decode_compute_size.restype = C.c_size_t
        L.bra_rle_decode_compute_size.argtypes = [u8p, C.c_size_t]
        L.bra_huffman_encode.restype = C.POINTER(_HuffChunk)
        L.bra_huffman_encode.argtypes = [u8p, C.c_uint32]
        L.bra_huffman_decode.restype = C.c_void_p
        L.bra_huffman_decode.argtypes = [C.POINTER(_HuffMeta), u8p, u32p]
        L.bra_huffman_chunk_free.restype = None
        L.bra_huffman_chunk_free.argtypes = [C.POINTER(_HuffChunk)]

    def _take(self, ptr, n):
        if not ptr:
            return None
        out = C.string_at(ptr, n)
        self.libc.free(ptr)
        return out

    def crc32c(self, data, prev=0, impl="bra_crc32c"):
        p, keep, n = _buf(data)
        return int(getattr(self.lib, impl)(C.cast(p, C.c_void_p), n, prev))

    def crc32c_combine(self, a, b, len_b):
        return int(self.lib.bra_crc32c_combine(a, b, len_b & 0xFFFFFFFF))

    def bwt_encode(self, data):
        p, keep, n = _buf(data)
        pi = C.c_uint32(0)
        return self._take(self.lib.bra_bwt_encode(p, n, C.byref(pi)), n), int(pi.value)

    def bwt_encode2(self, data):
        p, keep, n = _buf(data)
        pi = C.c_uint32(0)
        out = np.zeros(n, dtype=np.uint8)
        ok = self.lib.bra_bwt_encode2(p, n, C.byref(pi), out.ctypes.data_as(u8p))
        return (out.tobytes(), int(pi.value)) if ok else (None, None)

    def bwt_decode(self, data, primary):
        p, keep, n = _buf(data)
        return self._take(self.lib.bra_bwt_decode(p, n, primary), n)

    def bwt_decode2(self, data, primary):
        p, keep, n = _buf(data)
        tr = np.zeros(n, dtype=np.uint32)
        out = np.zeros(n, dtype=np.uint8)
        self.lib.bra_bwt_decode2(p, n, primary, tr.ctypes.data_as(u32p), out.ctypes.data_as(u8p))
        return out.tobytes()

    def mtf_encode(self, data):
        p, keep, n = _buf(data)
        return self._take(self.lib.bra_mtf_encode(p, n), n)

    def mtf_decode(self, data):
        p, keep, n = _buf(data)
        return self._take(self.lib.bra_mtf_decode(p, n), n)

    def rle_encode(self, data):
        p, keep, n = _buf(data)
        out = C.c_void_p(None)
        on = C.c_size_t(0)
        if not self.lib.bra_rle_encode(p, n, C.byref(out), C.byref(on)):
            return None
        return self._take(out.value, on.value)

    def rle_decode_size(self, data):
        p, keep, n = _buf(data)
        return int(self.lib.bra_rle_decode_compute_size(p, n))

    def rle_decode(self, data):
        p, keep, n = _buf(data)
        out = C.c_void_p(None)
        on = C.c_size_t(0)
        if not self.lib.bra_rle_decode(p, n, C.byref(out), C.byref(on)):
            return None
        return self._take(out.value, on.value)

    def huffman_encode(self, data):
        p, keep, n = _buf(data)
        ch = self.lib.bra_huffman_encode(p, n)
        if not ch:
            return None
        meta = ch.contents.meta
        lengths = bytes(meta.lengths)
        assert meta.orig_size == n
        payload = C.string_at(ch.contents.data, meta.encoded_size)
        self.lib.bra_huffman_chunk_free(ch)
        return lengths, payload

    def huffman_decode(self, lengths, payload, orig_size, encoded_size=None):
        meta = _HuffMeta()
        C.memmove(meta.lengths, bytes(lengths), 256)
        meta.orig_size = orig_size
        p, keep, n = _buf(payload)
        meta.encoded_size = n if encoded_size is None else encoded_size
        on = C.c_uint32(0)
        ptr = self.lib.bra_huffman_decode(C.byref(meta), p, C.byref(on))
        if not ptr:
            return None
        return self._take(ptr, on.value)

    # whole chain, looping the stage API exactly like reference chunks.c:214-246 / :362-397
    def encode_block(self, data):
        crc = self.crc32c(data)
        l, pi = self.bwt_encode2(data)
        m = self.mtf_encode(l)
        r = self.rle_encode(m)
        lengths, payload = self.huffman_encode(r)
        hdr = pi.to_bytes(4, "little") + lengths + len(r).to_bytes(4, "little") + len(payload).to_bytes(4, "little")
        return hdr, payload, crc

    def decode_block(self, hdr, payload):
        pi = int.from_bytes(hdr[0:4], "little")
        rn = int.from_bytes(hdr[260:264], "little")
        r = self.huffman_decode(hdr[4:260], payload, rn)
        if r is None:
            return None
        m = self.rle_decode(r)
        if m is None or pi >= len(m):
            return None
        return self.bwt_decode2(self.mtf_decode(m), pi)


def have_ref():
    return os.path.exists(REF_SO)


def load_ref():
    return RefApi(REF_SO)
