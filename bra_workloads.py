"""Deterministic synthetic workloads of SURVEY.md section 8(d), as numpy uint8 arrays in host memory.

Measurement / test tooling, deliberately outside the product library: tools/bra_gen.c -> tools/libbra_gen.so
(built on first use and by __graft_entry__.build()). bench.py's reference arm uses these without ever mapping
libbra_b200.so.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(ROOT, "tools", "bra_gen.c")
SO = os.path.join(ROOT, "tools", "libbra_gen.so")

_lib = None


def build():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c17", "-Wall", "-Wextra", "-o", SO, SRC])
    return SO


def _l():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.bra_gen_random.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.bra_gen_text.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_char_p), C.c_uint32]
        L.bra_gen_periodic.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
        _lib = L
    return _lib


def gen_random(n: int, seed: int = 2):
    """C3: bytes of successive splitmix64(seed) outputs, little-endian."""
    import numpy as np
    out = np.empty(n, dtype=np.uint8)
    _l().bra_gen_random(out.ctypes.data, n, seed)
    return out


def gen_text(n: int, vocab, seed: int = 1):
    """C2: random words of `vocab` joined by single spaces."""
    import numpy as np
    out = np.empty(n, dtype=np.uint8)
    arr = (C.c_char_p * len(vocab))(*[v.encode("latin-1") for v in vocab])
    _l().bra_gen_text(out.ctypes.data, n, seed, arr, len(vocab))
    return out


def gen_periodic(n: int, pattern: bytes):
    """C4a / C4b: `pattern` repeated and truncated to n bytes."""
    import numpy as np
    out = np.empty(n, dtype=np.uint8)
    pat = np.frombuffer(pattern, dtype=np.uint8)
    _l().bra_gen_periodic(out.ctypes.data, n, pat.ctypes.data, len(pattern))
    return out


HEX16 = b"0123456789abcdef"


def repeat251_pattern():
    """The 251-byte pattern of C4b (splitmix64 seed 4)."""
    return gen_random(251, 4).tobytes()


def make(kind: str, nbytes: int, seed: int, vocab=None):
    """One of the BASELINE shapes by name: text (C2), random (C3), periodic (C4a), repeat251 (C4b)."""
    if kind == "text":
        return gen_text(nbytes, vocab, seed)
    if kind == "random":
        return gen_random(nbytes, seed + 1)
    if kind == "periodic":
        return gen_periodic(nbytes, HEX16)
    if kind == "repeat251":
        return gen_periodic(nbytes, repeat251_pattern())
    raise ValueError(f"unknown workload {kind}")
