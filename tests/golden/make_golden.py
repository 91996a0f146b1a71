#!/usr/bin/env python3
"""Generate tests/golden/reference_vectors.json from the REFERENCE ITSELF.

Runs only in the dev container: it needs /root/reference (for the two text fixtures) and
oracle/_ref/libbra_ref.so (the reference's own encoder/CRC sources compiled in place by
oracle/Makefile). The JSON it writes is committed, so that the GPU box -- which has no
/root/reference -- can check both the oracle and the CUDA path against reference outputs.

Sections
  unit_tests : the known-answer vectors of the reference's own tests
               (reference test/test_bra_encoders.cpp:23-402, test/test_bra_crc32c.cpp:17-135),
               transcribed as input/expected pairs and re-verified here against the compiled
               reference before being written.
  blocks     : whole-chain outputs of the compiled reference for small inputs (fixtures, BWT tie
               cases, run-length edge cases, seeded random data): every intermediate stage as hex.
  vocab      : whitespace-split tokens of the reference's lorem fixture, the vocabulary of the
               "English-like" synthetic workload.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import load_ref  # noqa: E402

REF = "/root/reference"


def hx(b):
    return bytes(b).hex()


def main():
    ref = load_ref()
    test_txt = open(os.path.join(REF, "test/test.txt"), "rb").read()
    lorem = open(os.path.join(REF, "test/fixtures/lorem.txt"), "rb").read()

    unit = {}
    # --- RLE (test_bra_encoders.cpp:23-114)
    unit["rle"] = [
        {"in": hx(b"A" * 10), "out": hx(bytes([0xF7]) + b"A")},
        {"in": hx(b"AAAAABBBCD"), "out": hx(bytes([0xFC]) + b"A" + bytes([0xFE]) + b"B" + bytes([0x01]) + b"CD")},
        {"in": hx(b"ABCDEFGH"), "out": hx(bytes([0x07]) + b"ABCDEFGH")},
    ]
    for v in unit["rle"]:
        assert ref.rle_encode(bytes.fromhex(v["in"])) == bytes.fromhex(v["out"])
        assert ref.rle_decode(bytes.fromhex(v["out"])) == bytes.fromhex(v["in"])
    # --- BWT (:134-150)
    unit["bwt"] = [
        {"in": hx(b"BANANA"), "out": hx(b"NNBAAA"), "primary": 3},
        {"in": hx(b"The quick brown fox jumps over the lazy dog."),
         "out": hx(b"kynxeserg.l i hhv otTu c uwd rfm ebp qjoooza"), "primary": 9},
    ]
    for v in unit["bwt"]:
        assert ref.bwt_encode(bytes.fromhex(v["in"])) == (bytes.fromhex(v["out"]), v["primary"])
        assert ref.bwt_decode(bytes.fromhex(v["out"]), v["primary"]) == bytes.fromhex(v["in"])
    # --- MTF (:152-170, :199-218)
    unit["mtf"] = [
        {"in": hx(b"BANANA"), "out": hx(bytes([ord("B"), ord("B"), ord("N"), 1, 1, 1]))},
        {"in": hx(b"NNBAAA"), "out": hx(bytes([0x4E, 0x00, 0x43, 0x43, 0x00, 0x00]))},
    ]
    for v in unit["mtf"]:
        assert ref.mtf_encode(bytes.fromhex(v["in"])) == bytes.fromhex(v["out"])
        assert ref.mtf_decode(bytes.fromhex(v["out"])) == bytes.fromhex(v["in"])
    # --- Huffman (:262-365)
    unit["huffman"] = []
    for data in (b"BANANA", b"AAAAA", b"AAAAAAAA", bytes([0x4E, 0x00, 0x43, 0x43, 0x00, 0x00])):
        lengths, payload = ref.huffman_encode(data)
        unit["huffman"].append({"in": hx(data), "lengths": hx(lengths), "payload": hx(payload)})
    lengths, payload = ref.huffman_encode(b"BANANA")
    assert (lengths[ord("A")], lengths[ord("B")], lengths[ord("N")]) == (1, 2, 2) and payload == bytes([155, 0])
    assert ref.huffman_encode(b"AAAAA")[1] == b"\x00" and ref.huffman_encode(b"AAAAAAAA")[1] == b"\x00"
    assert ref.huffman_encode(b"") is None
    # --- CRC32C (test_bra_crc32c.cpp:17-135)
    unit["crc32c"] = [
        {"in": hx(b"123456789"), "crc": 0xE3069283},
        {"in": hx(b"Hello World!"), "crc": 0xFE6CF1DC},
        {"in": hx(b""), "crc": 0},
        {"in": hx(b"The quick brown fox jumps over the lazy dog"), "crc": ref.crc32c(b"The quick brown fox jumps over the lazy dog")},
    ]
    for v in unit["crc32c"]:
        for impl in ("bra_crc32c", "bra_crc32c_table", "bra_crc32c_sse42"):
            assert ref.crc32c(bytes.fromhex(v["in"]), impl=impl) == v["crc"]

    # --- whole-chain blocks
    rng = random.Random(20261018)
    inputs = {
        "test_txt": test_txt,
        "lorem_txt": lorem,
        "abcabcabcabc": b"abcabcabcabc",
        "A_x8": b"A" * 8,
        "hex16_x256": b"0123456789abcdef" * 256,
        "tie_bcabcabca": b"bcabcabca",
        "tie_cabcabcab": b"cabcabcab",
        "tie_babababa": b"babababa",
        "tie_zzzy_x3": b"zzzyzzzyzzzy",
        "tie_bca_x2000": b"bca" * 2000,
        "single_byte": b"Q",
        "two_bytes": b"ba",
        "zeros_4096": bytes(4096),
        "run_edges": b"".join(bytes([i & 0xFF]) * L for i, L in enumerate([1, 2, 3, 127, 128, 129, 130, 131, 255, 256, 257, 258, 259, 383, 384, 385, 386, 1, 2, 300])),
        "literal_300": bytes((i * 7 + 3) & 0xFF for i in range(300)),
        "random_1500": bytes(rng.randrange(256) for _ in range(1500)),
        "random_ab_2000": bytes(rng.choice(b"ab") for _ in range(2000)),
        "all_256_x3": bytes(range(256)) * 3,
        "period251_x9": bytes(rng.randrange(256) for _ in range(251)) * 9,
    }
    blocks = {}
    for name, data in inputs.items():
        crc = ref.crc32c(data)
        l, pi = ref.bwt_encode2(data)
        m = ref.mtf_encode(l)
        r = ref.rle_encode(m)
        lengths, payload = ref.huffman_encode(r)
        assert ref.decode_block(pi.to_bytes(4, "little") + lengths + len(r).to_bytes(4, "little") + len(payload).to_bytes(4, "little"), payload) == data
        blocks[name] = {
            "in": hx(data), "crc32c": crc, "primary": pi, "bwt": hx(l), "mtf": hx(m), "rle": hx(r),
            "lengths": hx(lengths), "payload": hx(payload),
            "crc_lengths": ref.crc32c(lengths), "crc_payload": ref.crc32c(payload),
        }
    # SURVEY.md section 8(c) probe table -- re-derived here, must still hold
    b = blocks["test_txt"]
    assert (len(test_txt), b["crc32c"], b["primary"], len(b["rle"]) // 2, len(b["payload"]) // 2) == (19, 0x1949C18E, 6, 20, 11)
    b = blocks["lorem_txt"]
    assert (len(lorem), b["crc32c"], b["primary"], len(b["rle"]) // 2, len(b["payload"]) // 2) == (3039, 0xB3AB946D, 577, 2638, 1325)
    assert blocks["tie_bcabcabca"]["primary"] == 3 and blocks["tie_cabcabcab"]["primary"] == 6
    assert blocks["tie_babababa"]["primary"] == 4 and blocks["tie_zzzy_x3"]["primary"] == 9
    assert blocks["tie_bca_x2000"]["primary"] == 2000

    vocab = [t.decode("latin-1") for t in lorem.split()]
    out = {"generated_by": "tests/golden/make_golden.py against oracle/_ref/libbra_ref.so (reference sources compiled in place)",
           "unit_tests": unit, "blocks": blocks, "vocab": vocab}
    path = os.path.join(HERE, "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes;", len(blocks), "blocks;", len(vocab), "vocab tokens")


if __name__ == "__main__":
    main()
