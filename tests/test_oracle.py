"""CPU: pins the oracle (oracle/bra_oracle.c) to the reference.

1. against the golden vectors of the reference's own unit tests and the whole-chain outputs the
   compiled reference produced (tests/golden/reference_vectors.json, made by make_golden.py);
2. against oracle/_ref/libbra_ref.so (the reference's own sources compiled in place) on seeded
   random inputs, when that library is present (dev container and, prebuilt, the GPU box).
"""
import random

import numpy as np
import pytest

import bra_workloads as wl

from oracle_lib import have_ref, load_ref

H = bytes.fromhex


def test_unit_vectors_rle(golden, oracle):
    for v in golden["unit_tests"]["rle"]:
        assert oracle.rle_encode(H(v["in"])) == H(v["out"])
        assert oracle.rle_decode(H(v["out"])) == H(v["in"])
        assert oracle.rle_decode_size(H(v["out"])) == len(H(v["in"]))


def test_unit_vectors_bwt(golden, oracle):
    for v in golden["unit_tests"]["bwt"]:
        for naive in (False, True):
            assert oracle.bwt_encode(H(v["in"]), naive=naive) == (H(v["out"]), v["primary"])
        assert oracle.bwt_decode(H(v["out"]), v["primary"]) == H(v["in"])


def test_unit_vectors_mtf(golden, oracle):
    for v in golden["unit_tests"]["mtf"]:
        assert oracle.mtf_encode(H(v["in"])) == H(v["out"])
        assert oracle.mtf_decode(H(v["out"])) == H(v["in"])


def test_unit_vectors_huffman(golden, oracle):
    for v in golden["unit_tests"]["huffman"]:
        lengths, payload = oracle.huffman_encode(H(v["in"]))
        assert lengths == H(v["lengths"]) and payload == H(v["payload"])
        assert oracle.huffman_decode(lengths, payload, len(H(v["in"]))) == H(v["in"])
    assert oracle.huffman_encode(b"") is None  # reference test_bra_encoders.cpp:358-365


def test_unit_vectors_crc(golden, oracle):
    for v in golden["unit_tests"]["crc32c"]:
        assert oracle.crc32c(H(v["in"])) == v["crc"]
    d = b"123456789"
    assert oracle.crc32c(d[5:], oracle.crc32c(d[:5])) == 0xE3069283  # incremental, test_bra_crc32c.cpp:27-31
    d = b"Hello World!"
    assert oracle.crc32c_combine(oracle.crc32c(d[:6]), oracle.crc32c(d[6:]), 6) == 0xFE6CF1DC


def test_golden_blocks(golden, oracle):
    for name, b in golden["blocks"].items():
        data = H(b["in"])
        assert oracle.crc32c(data) == b["crc32c"], name
        use_naive = len(data) <= 4096 and not name.startswith(("zeros", "tie_bca_x2000", "hex16"))
        l, pi = oracle.bwt_encode(data)
        assert (l, pi) == (H(b["bwt"]), b["primary"]), name
        if use_naive:
            assert oracle.bwt_encode(data, naive=True) == (l, pi), name
        m = oracle.mtf_encode(l)
        assert m == H(b["mtf"]), name
        r = oracle.rle_encode(m)
        assert r == H(b["rle"]), name
        lengths, payload = oracle.huffman_encode(r)
        assert lengths == H(b["lengths"]) and payload == H(b["payload"]), name
        hdr, pay, crc = oracle.encode_block(data)
        assert pay == payload and crc == b["crc32c"] and hdr[4:260] == lengths
        assert int.from_bytes(hdr[:4], "little") == pi
        assert oracle.decode_block(hdr, pay, len(data)) == data, name


def _gen(rng, kind, n):
    if kind == 0:
        return bytes(rng.randrange(256) for _ in range(n))
    if kind == 1:
        return bytes(rng.choice(b"ab") for _ in range(n))
    if kind == 2:
        out = bytearray()
        while len(out) < n:
            out += bytes([rng.randrange(4)]) * rng.choice([1, 1, 2, 2, 3, 4, 5, 127, 128, 129, 130, 131, 255, 256, 257, 300, 390])
        return bytes(out[:n])
    p = bytes(rng.randrange(3) for _ in range(rng.choice([1, 2, 3, 4, 5, 7, 16])))
    return (p * (n // len(p) + 1))[:n] if kind == 3 else p * max(1, n // len(p))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libbra_ref.so not built (needs /root/reference)")
def test_oracle_vs_compiled_reference(oracle):
    ref = load_ref()
    rng = random.Random(7)
    for it in range(400):
        n = rng.choice([1, 2, 3, 5, 8, 17, 64, 200, 513, 1000])
        d = _gen(rng, it % 5, n)
        if not d:
            continue
        assert oracle.encode_block(d) == ref.encode_block(d)
        assert oracle.encode_block(d, naive_bwt=True) == ref.encode_block(d)
        hdr, pay, _ = oracle.encode_block(d)
        assert oracle.decode_block(hdr, pay, len(d)) == d == ref.decode_block(hdr, pay)
        k = rng.randrange(len(d) + 1)
        assert oracle.crc32c_combine(oracle.crc32c(d[:k]), oracle.crc32c(d[k:]), len(d) - k) == ref.crc32c(d)
        for impl in ("bra_crc32c", "bra_crc32c_table", "bra_crc32c_sse42"):
            assert ref.crc32c(d, impl=impl) == oracle.crc32c(d)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libbra_ref.so not built (needs /root/reference)")
def test_oracle_error_paths_vs_reference(oracle, capfd):
    ref = load_ref()
    for bad in (b"\x05abc", b"\xfe", b"\x80", b"\x80\x80", b"\x00", b"\x00a\x80\xffb"):
        assert oracle.rle_decode_size(bad) == ref.rle_decode_size(bad)
        assert oracle.rle_decode(bad) == ref.rle_decode(bad)
    rng = random.Random(3)
    lengths, payload = oracle.huffman_encode(b"BANANA")
    for _ in range(300):
        l = bytearray(lengths)
        p = bytearray(payload + bytes(rng.randrange(256) for _ in range(rng.randrange(3))))
        for _ in range(rng.randrange(3)):
            l[rng.choice([65, 66, 78, rng.randrange(256)])] = rng.randrange(0, 6)
        if p and rng.random() < 0.5:
            p[rng.randrange(len(p))] ^= 1 << rng.randrange(8)
        osz = rng.choice([6, 6, 5, 7, 3])
        assert oracle.huffman_decode(bytes(l), bytes(p), osz) == ref.huffman_decode(bytes(l), bytes(p), osz)
    capfd.readouterr()


def test_fast_bwt_matches_naive_on_larger_inputs(oracle):
    rng = np.random.default_rng(11)
    for n in (4096, 20000):
        d = rng.integers(0, 4, n, dtype=np.uint8).tobytes()
        assert oracle.bwt_encode(d) == oracle.bwt_encode(d, naive=True)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libbra_ref.so not built (needs /root/reference)")
def test_fast_bwt_pinned_against_compiled_reference_at_block_scale(oracle, pkg, vocab):
    """The oracle's O(n log^2 n) BWT is what every large GPU comparison leans on (the reference's own rotation sort,
    bra_bwt.c:73-108, is O(n^2 log n) on repeats). Pin it against the compiled reference on the BASELINE shapes at the
    largest sizes the reference finishes in seconds: 64 KiB of text and random bytes, 16 KiB of period-16 data (C4a:
    n/16-way ties, primary 0) and of a 251-byte pattern repeated (C4b: every rotation distinct, LCP ~ n), plus the
    whole chain on top of it."""
    ref = load_ref()
    rnd251 = wl.gen_random(251, 4).tobytes()
    cases = {
        "text64k": wl.gen_text(65536, vocab, 3).tobytes(),
        "random64k": wl.gen_random(65536, 5).tobytes(),
        "period16_16k": b"0123456789abcdef" * 1024,
        "repeat251_16k": (rnd251 * 70)[:16384],
        "repeat251_ragged": (rnd251 * 30)[:7001],
        "runs24k": (b"a" * 300 + b"b" * 129 + b"ab" * 64) * 44,
    }
    for name, d in cases.items():
        exp = ref.bwt_encode2(d)
        assert oracle.bwt_encode(d) == exp, name
        assert oracle.encode_block(d) == ref.encode_block(d), name
    assert ref.bwt_encode2(cases["period16_16k"])[1] == 0
