#!/usr/bin/env python3
"""Generate tests/golden/reference_archives.json with the UNMODIFIED reference CLI
(oracle/_ref/bra_ref = every reference source compiled by oracle/Makefile:ref_cli, no cmake).

For each input: `bra_ref -c -o <name>.BRa <name>` -> archive bytes (hex for the small one, size + sha256 +
trailing entry CRC for the larger ones). The inputs are reproducible on any box: the lorem fixture (hex in
reference_vectors.json) and outputs of the library's deterministic generators.
Dev container only (needs oracle/_ref/bra_ref and the built library for the generators)."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bra_pkg  # noqa: E402
import bra_workloads as wl  # noqa: E402


def inputs(pkg, golden):
    vocab = golden["vocab"]
    return {
        "lorem.txt": bytes.fromhex(golden["blocks"]["lorem_txt"]["in"]),                 # 1 chunk (reference test_bra.cpp:353-398)
        "text700k.txt": wl.gen_text(700_000, vocab, 21).tobytes(),                      # 3 chunks of 256 KiB, ragged tail
        # text with an incompressible stretch inside one chunk (long runs are avoided on purpose: the reference's
        # O(n^2 log n) rotation sort needs hours on them, reference src/encoders/bra_bwt.h:27-29)
        "mixed600k.bin": wl.gen_text(400_000, vocab, 22).tobytes() + wl.gen_random(50_000, 23).tobytes() + wl.gen_text(150_000, vocab, 24).tobytes(),
    }


def main():
    pkg = bra_pkg.load()
    golden = json.load(open(os.path.join(HERE, "reference_vectors.json")))
    bra = os.path.join(ROOT, "oracle", "_ref", "bra_ref")
    out = {"generated_by": "tests/golden/make_golden_archives.py with oracle/_ref/bra_ref (unmodified reference CLI)", "archives": {}}
    with tempfile.TemporaryDirectory() as d:
        for name, data in inputs(pkg, golden).items():
            open(os.path.join(d, name), "wb").write(data)
            arc = name + ".BRa"
            subprocess.run([bra, "-c", "-o", arc, name], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            a = open(os.path.join(d, arc), "rb").read()
            e = {"input_sha256": hashlib.sha256(data).hexdigest(), "input_size": len(data), "size": len(a),
                 "sha256": hashlib.sha256(a).hexdigest(), "entry_crc32c": int.from_bytes(a[-4:], "little")}
            if len(a) < 4096:
                e["hex"] = a.hex()
            out["archives"][name] = e
            print(name, len(data), "->", len(a), hex(e["entry_crc32c"]))
    assert out["archives"]["lorem.txt"]["size"] == 1623 and out["archives"]["lorem.txt"]["entry_crc32c"] == 0x74F1DA57
    json.dump(out, open(os.path.join(HERE, "reference_archives.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
