"""CPU: the scalar logic the kernels share with the host (br-archive_b200/csrc/bra_hd.h, built as
libbra_hostlogic.so) against the oracle: Huffman tree replay, canonical codes, decode tables,
GF(2) CRC folding. This is the exact code one GPU thread per block executes."""
import ctypes as C
import random

import numpy as np
import pytest

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def hl(pkg):
    import os
    if not os.path.exists(pkg.HOSTLOGIC_PATH):
        pkg.build()
    L = C.CDLL(pkg.HOSTLOGIC_PATH)
    L.hl_gf_mul.restype = C.c_uint32
    L.hl_gf_mul.argtypes = [C.c_uint32, C.c_uint32]
    L.hl_crc_combine.restype = C.c_uint32
    L.hl_crc_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    L.hl_huf_lengths.restype = C.c_uint32
    L.hl_huf_lengths.argtypes = [u32p, u8p]
    L.hl_huf_canonical.argtypes = [u8p, u32p]
    L.hl_huf_decode.restype = C.c_int
    L.hl_huf_decode.argtypes = [u8p, u8p, C.c_uint32, C.c_uint32, u8p]
    return L


def _lengths(hl, freq):
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    out = np.zeros(256, dtype=np.uint8)
    k = hl.hl_huf_lengths(f.ctypes.data_as(u32p), out.ctypes.data_as(u8p))
    return k, out


def test_huffman_lengths_replay(hl, oracle):
    rng = random.Random(5)
    for it in range(1500):
        k = rng.choice([1, 2, 3, 5, 17, 100, 256])
        freq = np.zeros(256, dtype=np.uint32)
        for s in rng.sample(range(256), k):
            freq[s] = [rng.randrange(1, 4), rng.randrange(1, 1000), 1 << rng.randrange(20), rng.randrange(1, 1 << 24)][it % 4]
        n, got = _lengths(hl, freq)
        assert n == k
        assert bytes(got) == bytes(oracle.huffman_lengths(freq)), (it, freq[freq > 0])
    assert _lengths(hl, np.zeros(256, dtype=np.uint32))[0] == 0


def test_huffman_lengths_fibonacci_depth(hl, oracle):
    fib = [1, 1]
    while len(fib) < 33:
        fib.append(fib[-1] + fib[-2])
    freq = np.zeros(256, dtype=np.uint32)
    freq[10:10 + len(fib)] = fib
    _, got = _lengths(hl, freq)
    assert bytes(got) == bytes(oracle.huffman_lengths(freq))
    assert got.max() == 32


def test_canonical_codes(hl, oracle):
    rng = random.Random(9)
    for _ in range(300):
        lengths = np.zeros(256, dtype=np.uint8)
        for s in rng.sample(range(256), rng.choice([1, 2, 40, 256])):
            lengths[s] = rng.randrange(1, 40)
        got = np.zeros(256, dtype=np.uint32)
        hl.hl_huf_canonical(lengths.ctypes.data_as(u8p), got.ctypes.data_as(u32p))
        assert (got == oracle.huffman_canonical(lengths)).all()


def test_decode_tables_roundtrip(hl, oracle):
    rng = np.random.default_rng(3)
    for it in range(60):
        n = int(rng.integers(1, 5000))
        k = int(rng.choice([1, 2, 3, 16, 256]))
        p = rng.dirichlet(np.ones(k) * (0.05 if it % 2 else 1.0))
        data = rng.choice(k, size=n, p=p).astype(np.uint8)
        lengths, payload = oracle.huffman_encode(data.tobytes())
        l = np.frombuffer(lengths, dtype=np.uint8).copy()
        pay = np.frombuffer(payload, dtype=np.uint8).copy()
        out = np.zeros(n, dtype=np.uint8)
        rc = hl.hl_huf_decode(l.ctypes.data_as(u8p), pay.ctypes.data_as(u8p), len(pay), n, out.ctypes.data_as(u8p))
        assert rc == 0 and out.tobytes() == data.tobytes()


def test_decode_tables_reject_oversubscribed(hl):
    l = np.zeros(256, dtype=np.uint8)
    l[0:3] = 1  # three 1-bit codes cannot exist
    out = np.zeros(4, dtype=np.uint8)
    pay = np.zeros(4, dtype=np.uint8)
    assert hl.hl_huf_decode(l.ctypes.data_as(u8p), pay.ctypes.data_as(u8p), 4, 4, out.ctypes.data_as(u8p)) == -1
    l[:] = 0
    l[7] = 33  # longer than the decoder supports
    assert hl.hl_huf_decode(l.ctypes.data_as(u8p), pay.ctypes.data_as(u8p), 4, 1, out.ctypes.data_as(u8p)) == -1


def test_crc_fold(hl, oracle):
    rng = random.Random(1)
    for _ in range(300):
        d = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 400)))
        k = rng.randrange(len(d) + 1)
        a, b = oracle.crc32c(d[:k]), oracle.crc32c(d[k:])
        assert hl.hl_crc_combine(a, b, len(d) - k) == oracle.crc32c(d) == oracle.crc32c_combine(a, b, len(d) - k)
    # long zero runs: combine with huge lengths must agree with the oracle's square-and-multiply
    for ln in (1 << 20, (1 << 32) - 1, 123456789):
        assert hl.hl_crc_combine(0x12345678, 0x9ABCDEF0, ln) == oracle.crc32c_combine(0x12345678, 0x9ABCDEF0, ln)


def test_stage_plan_covers_every_block(hl):
    """bra_stage_plan (host-path pipeline): stages are non-empty, never wider than the context's batch, and add up to the job."""
    hl.hl_stage_plan.restype = C.c_uint32
    hl.hl_stage_plan.argtypes = [C.c_uint64, C.c_uint32, u32p]
    for hb in (1, 2, 3, 4, 8, 16, 64, 256, 1024, 32768):
        for n in list(range(1, 700)) + [1023, 1024, 1025, 4096, 5000, 65536, 100003]:
            plan = np.zeros(n // hb + 4, dtype=np.uint32)
            k = hl.hl_stage_plan(n, hb, plan.ctypes.data_as(u32p))
            assert 1 <= k <= len(plan)
            p = plan[:k]
            assert p.min() >= 1 and p.max() <= hb and int(p.sum()) == n, (n, hb, p.tolist())
    plan = np.zeros(8, dtype=np.uint32)
    k = hl.hl_stage_plan(1024, 1024, plan.ctypes.data_as(u32p))
    assert plan[:k].tolist() == [102, 602, 256, 64]  # short head, wide middle, shrinking tail
    # encoding: short head, then stages as wide as the context allows (output copies are short, nothing to taper)
    hl.hl_stage_plan_encode.restype = C.c_uint32
    hl.hl_stage_plan_encode.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, u32p]
    for hb in (1, 2, 3, 4, 8, 16, 64, 256, 1024, 32768):
        for n in list(range(1, 700)) + [1023, 1024, 1025, 4096, 5000, 65536, 100003]:
            for div in (0, 2, 8, 16):
                plan = np.zeros(n // hb + 4, dtype=np.uint32)
                k = hl.hl_stage_plan_encode(n, hb, div, plan.ctypes.data_as(u32p))
                assert 1 <= k <= len(plan)
                p = plan[:k]
                assert p.min() >= 1 and p.max() <= hb and int(p.sum()) == n, (n, hb, div, p.tolist())
    plan = np.zeros(8, dtype=np.uint32)
    k = hl.hl_stage_plan_encode(1024, 1024, 8, plan.ctypes.data_as(u32p))
    assert plan[:k].tolist() == [128, 896]


def test_finisher_rotation_compare_is_exact_to_the_byte(hl):
    """bra_rot_cmp_window (the BWT finisher's comparison, same code on the GPU): against a plain cyclic byte comparison of
    exactly `depth` bytes -- windows that wrap around the block end included. Looking even one byte deeper for some pairs
    makes 'equal' non-transitive (the bug tests/golden/fuzz_case_31337_608.json records)."""
    import json
    import os
    hl.hl_rot_cmp.restype = C.c_int
    hl.hl_rot_cmp.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]

    def check(T, pairs, frm, depth=64):
        p = len(T)
        buf = np.zeros(p + 16, dtype=np.uint8)
        buf[:p] = T
        buf[p:] = 0xEE  # bytes past the block must never matter
        ptr = buf.ctypes.data_as(u8p)
        idx = np.arange(depth)
        for a, c in pairs:
            wa, wc = bytes(T[(a + frm + idx) % p]), bytes(T[(c + frm + idx) % p])
            exp = (wa > wc) - (wa < wc)
            assert hl.hl_rot_cmp(ptr, p, a, c, frm % p, depth) == exp, (p, a, c, frm)

    rng = np.random.default_rng(5)
    for it in range(300):
        p = int(rng.integers(70, 700))
        # runs over a tiny alphabet: long ties that end just past the window
        T = np.repeat(rng.integers(0, 3, 64, dtype=np.uint8), rng.integers(1, 80, 64))[:p]
        if len(T) < p:
            T = np.resize(T, p)
        pairs = [(int(rng.integers(0, p)), int(rng.integers(0, p))) for _ in range(60)]
        check(T, pairs, int(rng.integers(0, 4 * p)))
    case = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_case_31337_608.json")))
    T = np.frombuffer(bytes.fromhex(case["data"])[case["block"]:], dtype=np.uint8)
    members = list(range(5578, 5833))
    check(T, [(a, c) for a in members[::3] for c in members[::5]], 4096)
    check(T, [(5766, 5633), (5766, 5637), (5633, 5766)], 4096)


def test_mtf_list_operations_match_the_oracle(hl, oracle):
    """bra_mtf_list_encode / bra_mtf_list_decode (the per-thread list the GPU replay kernel runs: sixteen entries per
    128-bit chunk, shift by byte permutes) against the oracle's move-to-front -- every rank, runs, front hits."""
    hl.hl_mtf.restype = None
    hl.hl_mtf.argtypes = [u8p, u8p, C.c_uint64, C.c_int]

    def run(data, decode):
        src = np.frombuffer(data, dtype=np.uint8).copy()
        dst = np.zeros(len(src), dtype=np.uint8)
        hl.hl_mtf(src.ctypes.data_as(u8p), dst.ctypes.data_as(u8p), len(src), decode)
        return dst.tobytes()

    rng = np.random.default_rng(11)
    cases = [rng.integers(0, 256, 20000, dtype=np.uint8).tobytes(),                      # uniform ranks, deep shifts
             np.repeat(rng.integers(0, 256, 500, dtype=np.uint8), rng.integers(1, 40, 500)).tobytes(),  # runs: rank 0
             rng.integers(0, 4, 5000, dtype=np.uint8).tobytes(),                         # small alphabet: ranks 0..3
             bytes(range(255, -1, -1)) * 8,                                              # always the last entry (rank 255)
             bytes([15, 16, 17, 31, 32, 0, 255, 254, 16, 15] * 50)]                      # chunk borders
    for data in cases:
        enc = run(data, 0)
        assert enc == oracle.mtf_encode(data)
        assert run(enc, 1) == data == oracle.mtf_decode(enc)
