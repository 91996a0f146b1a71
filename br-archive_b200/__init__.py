"""br-archive_b200 -- B200-native block-compression path of BR-Archive (CRC32C, BWT, MTF, RLE, Huffman).

The product is the C-ABI shared library ``libbra_b200.so`` built from ``csrc/`` (hand-written
sm_100a CUDA kernels, declared in ``include/*.h``). This Python module is only the thin ctypes
binding the tests and ``bench.py`` use; torch appears here purely for device memory and streams.

There is no CPU fallback: importing works without a GPU (so the CPU test-suite can check that
the library loads and exports its symbols), but every compute entry point needs a CUDA device,
and a missing ``libbra_b200.so`` raises immediately.

The directory name contains a hyphen, so import it through ``bra_pkg.load()`` (repo root),
which registers it as ``br_archive_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libbra_b200.so")
HOSTLOGIC_PATH = os.path.join(HERE, "libbra_hostlogic.so")

HDR_BYTES = 268
DISK_HDR_BYTES = 267

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)

# every symbol include/*.h declares; tests/test_abi.py checks the library exports them all
REFERENCE_API = [
    "bra_bwt_encode", "bra_bwt_encode2", "bra_bwt_decode", "bra_bwt_decode2",
    "bra_mtf_encode", "bra_mtf_encode2", "bra_mtf_decode", "bra_mtf_decode2",
    "bra_rle_encode", "bra_rle_decode_compute_size", "bra_rle_decode",
    "bra_huffman_encode", "bra_huffman_decode", "bra_huffman_chunk_free",
    "bra_crc32c", "bra_crc32c_table", "bra_crc32c_sse42", "bra_crc32c_combine", "bra_crc32c_use_sse42",
    "bra_init", "bra_quit", "bra_has_sse42",
]
BATCH_API = [
    "bra_b200_device_count", "bra_b200_ctx_create", "bra_b200_ctx_destroy", "bra_b200_block_size", "bra_b200_max_batch",
    "bra_b200_payload_stride", "bra_b200_workspace_bytes", "bra_b200_last_stats", "bra_b200_encode_device", "bra_b200_decode_device",
    "bra_b200_encode_bound", "bra_b200_encode_host", "bra_b200_decode_host", "bra_b200_list_host", "bra_b200_host_alloc", "bra_b200_host_free", "bra_b200_crc32c_submit", "bra_b200_crc32c_finish",
    "bra_b200_prof_enable", "bra_b200_prof_reset", "bra_b200_prof_count", "bra_b200_prof_read",
    "bra_b200_pool_create", "bra_b200_pool_destroy", "bra_b200_pool_workers", "bra_b200_pool_worker_ranges", "bra_b200_pool_encode_bound",
    "bra_b200_pool_encode_host", "bra_b200_pool_decode_host",
]


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libbra_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "-j8", "all"], stdout=out)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The product library. Raises if it has not been built: nothing here falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the compression path)")
    L = C.CDLL(LIB_PATH)
    L.bra_b200_device_count.restype = C.c_int
    L.bra_b200_ctx_create.restype = C.c_void_p
    L.bra_b200_ctx_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32]
    L.bra_b200_ctx_destroy.argtypes = [C.c_void_p]
    L.bra_b200_block_size.restype = C.c_uint32
    L.bra_b200_block_size.argtypes = [C.c_void_p]
    L.bra_b200_max_batch.restype = C.c_uint32
    L.bra_b200_max_batch.argtypes = [C.c_void_p]
    L.bra_b200_payload_stride.restype = C.c_uint64
    L.bra_b200_payload_stride.argtypes = [C.c_void_p]
    L.bra_b200_workspace_bytes.restype = C.c_uint64
    L.bra_b200_workspace_bytes.argtypes = [C.c_void_p]
    L.bra_b200_last_stats.argtypes = [C.c_void_p, u32p, u32p, C.POINTER(C.c_uint64)]
    L.bra_b200_encode_device.restype = C.c_int
    L.bra_b200_encode_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bra_b200_decode_device.restype = C.c_int
    L.bra_b200_decode_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
    L.bra_b200_encode_bound.restype = C.c_uint64
    L.bra_b200_encode_bound.argtypes = [C.c_void_p, C.c_uint64]
    L.bra_b200_encode_host.restype = C.c_int
    L.bra_b200_encode_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), u32p]
    L.bra_b200_decode_host.restype = C.c_int
    L.bra_b200_decode_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), u32p]
    L.bra_b200_list_host.restype = C.c_int
    L.bra_b200_list_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.bra_b200_crc32c_submit.restype = C.c_int
    L.bra_b200_crc32c_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.bra_b200_crc32c_finish.restype = C.c_int
    L.bra_b200_crc32c_finish.argtypes = [C.c_void_p, u32p]
    L.bra_b200_pool_create.restype = C.c_void_p
    L.bra_b200_pool_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_uint32, C.c_uint32, C.c_int]
    L.bra_b200_pool_destroy.argtypes = [C.c_void_p]
    L.bra_b200_pool_workers.restype = C.c_int
    L.bra_b200_pool_workers.argtypes = [C.c_void_p]
    L.bra_b200_pool_worker_ranges.restype = C.c_int
    L.bra_b200_pool_worker_ranges.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), u32p]
    L.bra_b200_pool_encode_bound.restype = C.c_uint64
    L.bra_b200_pool_encode_bound.argtypes = [C.c_void_p, C.c_uint64]
    L.bra_b200_pool_encode_host.restype = C.c_int
    L.bra_b200_pool_encode_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), u32p]
    L.bra_b200_pool_decode_host.restype = C.c_int
    L.bra_b200_pool_decode_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), u32p]
    L.bra_b200_prof_enable.argtypes = [C.c_int]
    L.bra_b200_prof_count.restype = C.c_int
    L.bra_b200_prof_read.restype = C.c_int
    L.bra_b200_prof_read.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    _lib = L
    return L


def prof_enable(timing: bool):
    lib().bra_b200_prof_enable(1 if timing else 0)


def prof_reset():
    lib().bra_b200_prof_reset()


def prof_read():
    """-> {kernel family: (launches, accumulated device ms)}; ms is 0 unless timing was enabled."""
    L = lib()
    out = {}
    for i in range(L.bra_b200_prof_count()):
        name, n, ms = C.c_char_p(), C.c_uint64(0), C.c_double(0.0)
        L.bra_b200_prof_read(i, C.byref(name), C.byref(n), C.byref(ms))
        out[name.value.decode()] = (int(n.value), float(ms.value))
    return out


# ---------------------------------------------------------------------------------------------
# multi-GPU host logic: blocks are independent, so ranks take contiguous block ranges and the per-rank
# CRC chains are folded on the host; no data-path collective exists.
# ---------------------------------------------------------------------------------------------
def shard_range(nblk: int, rank: int, world: int):
    """Contiguous block range [lo, hi) of `rank`; the first nblk % world ranks get one extra block."""
    base, extra = divmod(nblk, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def chain_bytes(block_lens):
    """Bytes the per-entry CRC chain covers for these blocks: 268-byte in-memory header + raw block each
    (reference src/io/lib_bra_io_file_chunks.c:248-249)."""
    return sum(HDR_BYTES + int(n) for n in block_lens)


def fold_crc_chains(parts):
    """parts: [(crc_chain_of_rank_started_from_0, bytes_covered)] in rank order -> CRC chain of the whole
    block sequence, via bra_crc32c_combine (host arithmetic, no GPU needed). Lengths above 2^32-1 are
    folded piecewise because the reference's combine takes a 32-bit length."""
    L = lib()
    L.bra_crc32c_combine.restype = C.c_uint32
    L.bra_crc32c_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    crc = 0
    for part_crc, nbytes in parts:
        if nbytes == 0:
            continue
        # advance crc over nbytes zero-effect bytes: combine(crc, 0-chain, len) in pieces, then xor in the part
        remaining = nbytes
        while remaining > 0xFFFFFFFF:
            crc = L.bra_crc32c_combine(crc, 0, 0xFFFFFFFF)
            remaining -= 0xFFFFFFFF
        crc = L.bra_crc32c_combine(crc, part_crc, remaining)
    return int(crc)


class Context:
    """One GPU context of the batched path (include/bra_b200.h). Tensors are torch CUDA uint8/int32."""

    def __init__(self, device: int = 0, block_size: int = 1 << 20, max_batch: int = 256):
        self.L = lib()
        self.device = device
        self.block_size = block_size
        self.handle = self.L.bra_b200_ctx_create(device, block_size, max_batch)
        if not self.handle:
            raise RuntimeError("bra_b200_ctx_create failed (no CUDA device, or out of device memory); there is no CPU fallback")
        self.payload_stride = int(self.L.bra_b200_payload_stride(self.handle))
        self.max_batch = int(self.L.bra_b200_max_batch(self.handle))

    def close(self):
        if self.handle:
            self.L.bra_b200_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        r, s, l = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        self.L.bra_b200_last_stats(self.handle, C.byref(r), C.byref(s), C.byref(l))
        return {"bwt_rounds": r.value, "huf_sweeps": s.value, "launches": l.value}

    # ---- device-resident -------------------------------------------------------------------
    def alloc_encode_outputs(self, nblk: int):
        import torch
        dev = torch.device("cuda", self.device)
        return (torch.empty(nblk * HDR_BYTES, dtype=torch.uint8, device=dev),
                torch.empty(nblk * self.payload_stride, dtype=torch.uint8, device=dev),
                torch.empty(nblk, dtype=torch.int32, device=dev))

    def alloc_decode_outputs(self, nblk: int):
        import torch
        dev = torch.device("cuda", self.device)
        return (torch.empty(nblk * self.block_size, dtype=torch.uint8, device=dev),
                torch.empty(nblk, dtype=torch.int32, device=dev),
                torch.empty(nblk, dtype=torch.int32, device=dev),
                torch.empty(nblk, dtype=torch.int32, device=dev))

    def encode_device(self, d_in, total: int, outputs=None, stream=None):
        """d_in: CUDA uint8 tensor holding `total` bytes. Returns (hdr, payload, crc_raw) tensors."""
        import torch
        nblk = (total + self.block_size - 1) // self.block_size
        last = total - (nblk - 1) * self.block_size
        if outputs is None:
            outputs = self.alloc_encode_outputs(nblk)
        hdr, pay, crc = outputs
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        rc = self.L.bra_b200_encode_device(self.handle, d_in.data_ptr(), nblk, last, hdr.data_ptr(), pay.data_ptr(), crc.data_ptr(), st)
        if rc != 0:
            raise RuntimeError(f"bra_b200_encode_device failed with code {rc}")
        return hdr, pay, crc

    def decode_device(self, hdr, pay, nblk: int, hint_r: int = 0, hint_c: int = 0, outputs=None, stream=None):
        import torch
        if outputs is None:
            outputs = self.alloc_decode_outputs(nblk)
        out, out_len, crc, status = outputs
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        rc = self.L.bra_b200_decode_device(self.handle, hdr.data_ptr(), pay.data_ptr(), nblk, hint_r, hint_c, out.data_ptr(),
                                           out_len.data_ptr(), crc.data_ptr(), status.data_ptr(), st)
        if rc != 0:
            raise RuntimeError(f"bra_b200_decode_device failed with code {rc}")
        return out, out_len, crc, status

    # ---- host buffers (numpy uint8 arrays or pinned torch CPU tensors) ------------------------------
    def encode_host(self, data, out=None, crc_chain: int = 0):
        import numpy as np
        n = int(data.nbytes) if hasattr(data, "nbytes") else int(data.numel())
        ptr = data.ctypes.data if hasattr(data, "ctypes") else data.data_ptr()
        bound = int(self.L.bra_b200_encode_bound(self.handle, n))
        if out is None:
            out = np.empty(bound, dtype=np.uint8)
        optr = out.ctypes.data if hasattr(out, "ctypes") else out.data_ptr()
        ocap = int(out.nbytes) if hasattr(out, "nbytes") else int(out.numel())
        osz = C.c_uint64(0)
        crc = C.c_uint32(crc_chain)
        rc = self.L.bra_b200_encode_host(self.handle, ptr, n, optr, ocap, C.byref(osz), C.byref(crc))
        if rc != 0:
            raise RuntimeError(f"bra_b200_encode_host failed with code {rc}")
        return out[: osz.value], int(crc.value)

    def decode_host(self, stream_bytes, out_cap: int, out=None, crc_chain: int = 0):
        import numpy as np
        n = int(stream_bytes.nbytes) if hasattr(stream_bytes, "nbytes") else int(stream_bytes.numel())
        ptr = stream_bytes.ctypes.data if hasattr(stream_bytes, "ctypes") else stream_bytes.data_ptr()
        if out is None:
            out = np.empty(max(out_cap, 1), dtype=np.uint8)
        optr = out.ctypes.data if hasattr(out, "ctypes") else out.data_ptr()
        osz = C.c_uint64(0)
        crc = C.c_uint32(crc_chain)
        rc = self.L.bra_b200_decode_host(self.handle, ptr, n, optr, out_cap, C.byref(osz), C.byref(crc))
        if rc != 0:
            raise RuntimeError(f"bra_b200_decode_host failed with code {rc}")
        return out[: osz.value], int(crc.value)

    def list_host(self, stream_bytes) -> int:
        """Plain size of a chunk stream, the reference's list mode (chunks.c:369-373): Huffman decode + RLE size pass only."""
        n = int(stream_bytes.nbytes) if hasattr(stream_bytes, "nbytes") else int(stream_bytes.numel())
        ptr = stream_bytes.ctypes.data if hasattr(stream_bytes, "ctypes") else stream_bytes.data_ptr()
        osz = C.c_uint64(0)
        rc = self.L.bra_b200_list_host(self.handle, ptr, n, C.byref(osz))
        if rc != 0:
            raise RuntimeError(f"bra_b200_list_host failed with code {rc}")
        return int(osz.value)


def _ptr_len(a):
    return (a.ctypes.data if hasattr(a, "ctypes") else a.data_ptr()), (int(a.nbytes) if hasattr(a, "nbytes") else int(a.numel()))


class Pool:
    """One job over several GPUs of this process (include/bra_b200.h, bra_b200_pool_*): the block list of one input is cut
    into ranges that per-GPU workers take from a shared queue; output is the ordered chunk stream and the folded CRC chain,
    byte-identical to a single context's."""

    def __init__(self, devices, block_size: int = 1 << 20, range_blocks: int = 64, workers_per_device: int = 2):
        self.L = lib()
        arr = (C.c_int * len(devices))(*devices)
        self.handle = self.L.bra_b200_pool_create(arr, len(devices), block_size, range_blocks, workers_per_device)
        if not self.handle:
            raise RuntimeError("bra_b200_pool_create failed (no CUDA device, or out of device memory); there is no CPU fallback")
        self.block_size = block_size

    def close(self):
        if self.handle:
            self.L.bra_b200_pool_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode_bound(self, total: int) -> int:
        return int(self.L.bra_b200_pool_encode_bound(self.handle, total))

    def stats(self):
        out = []
        for w in range(self.L.bra_b200_pool_workers(self.handle)):
            d, r = C.c_int(0), C.c_uint32(0)
            self.L.bra_b200_pool_worker_ranges(self.handle, w, C.byref(d), C.byref(r))
            out.append({"worker": w, "device": d.value, "ranges": r.value})
        return out

    def encode_host(self, data, out=None, crc_chain: int = 0):
        import numpy as np
        ptr, n = _ptr_len(data)
        if out is None:
            out = np.empty(self.encode_bound(n), dtype=np.uint8)
        optr, ocap = _ptr_len(out)
        osz, crc = C.c_uint64(0), C.c_uint32(crc_chain)
        rc = self.L.bra_b200_pool_encode_host(self.handle, ptr, n, optr, ocap, C.byref(osz), C.byref(crc))
        if rc != 0:
            raise RuntimeError(f"bra_b200_pool_encode_host failed with code {rc}")
        return out[: osz.value], int(crc.value)

    def decode_host(self, stream_bytes, out_cap: int, out=None, crc_chain: int = 0):
        import numpy as np
        ptr, n = _ptr_len(stream_bytes)
        if out is None:
            out = np.empty(max(out_cap, 1), dtype=np.uint8)
        optr, _ = _ptr_len(out)
        osz, crc = C.c_uint64(0), C.c_uint32(crc_chain)
        rc = self.L.bra_b200_pool_decode_host(self.handle, ptr, n, optr, out_cap, C.byref(osz), C.byref(crc))
        if rc != 0:
            raise RuntimeError(f"bra_b200_pool_decode_host failed with code {rc}")
        return out[: osz.value], int(crc.value)
