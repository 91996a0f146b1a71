"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(dirpath, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names.update(re.findall(r"\b(bra_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    declared = _declared_symbols()
    assert set(pkg.REFERENCE_API) | set(pkg.BATCH_API) == declared, declared ^ (set(pkg.REFERENCE_API) | set(pkg.BATCH_API))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by libbra_b200.so"


def test_no_cpu_fallback_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        return
    lib = pkg.lib()
    assert lib.bra_b200_device_count() <= 0
    assert not lib.bra_b200_ctx_create(0, 1 << 20, 4)  # fails loudly (NULL), no CPU path
    lib.bra_init.restype = C.c_bool
    assert lib.bra_init() is False


def test_product_does_not_touch_the_oracle():
    """Nothing under br-archive_b200/ or include/ may reference oracle/ (the checker is not the product)."""
    for base in ("br-archive_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".so", ".o")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "bra_oracle" not in text and "oracle_lib" not in text, os.path.join(dirpath, f)
