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


def trees(golden):
    """Multi-entry archives: {archive: (bra arguments, {relative path: bytes})}. `tree` is what the reference's own CLI tests
    pack (test/test_bra.cpp:310-351: `bra -c -r` over a directory, six empty files): directory entries first, then files in
    the order of reference src/prog/bra.cpp:337-358; `b.bin` is incompressible, so its entry is first compressed into the
    temporary file, found not smaller and rewritten STORED (reference chunks.c:268-278, meta_entries.c:194-205)."""
    vocab = golden["vocab"]
    tree = {"tree/lorem.txt": bytes.fromhex(golden["blocks"]["lorem_txt"]["in"]),
            "tree/sub/a.txt": wl.gen_text(300_000, vocab, 31).tobytes(),
            "tree/sub/b.bin": wl.gen_random(300_000, 32).tobytes(),
            "tree/sub/deep/c.txt": wl.gen_text(70_000, vocab, 33).tobytes()}
    for i in range(6):
        tree[f"tree/e{i}"] = b""
    return {"tree": (["-c", "-r", "-o", "tree.BRa", "tree"], tree),
            "stored_only": (["-c", "-o", "stored_only.BRa", "noise.bin"], {"noise.bin": wl.gen_random(600_000, 34).tobytes()})}


def write_tree(root, files):
    for rel, data in files.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(data)


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
    out["trees"] = {}
    for name, (argv, files) in trees(golden).items():
        with tempfile.TemporaryDirectory() as d:
            write_tree(d, files)
            subprocess.run([bra] + argv, cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            a = open(os.path.join(d, name + ".BRa"), "rb").read()
            out["trees"][name] = {"size": len(a), "sha256": hashlib.sha256(a).hexdigest(), "entries": int.from_bytes(a[4:8], "little"),
                                  "input_sha256": hashlib.sha256(b"".join(k.encode() + v for k, v in sorted(files.items()))).hexdigest()}
            print(name, "->", len(a), "bytes,", out["trees"][name]["entries"], "entries")
    assert out["archives"]["lorem.txt"]["size"] == 1623 and out["archives"]["lorem.txt"]["entry_crc32c"] == 0x74F1DA57
    json.dump(out, open(os.path.join(HERE, "reference_archives.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
