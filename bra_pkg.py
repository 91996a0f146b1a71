"""Loads the `br-archive_b200/` package (hyphenated directory) as module `br_archive_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load():
    if "br_archive_b200" in sys.modules:
        return sys.modules["br_archive_b200"]
    path = os.path.join(_ROOT, "br-archive_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("br_archive_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["br_archive_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
