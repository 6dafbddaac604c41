"""CPU: the C-ABI library builds, loads and exports every symbol include/spurfies_b200.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from spurfies_b200 import _lib
    header = open(os.path.join(ROOT, "include", "spurfies_b200.h")).read()
    names = sorted(set(re.findall(r"\b(spf_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(_lib.lib, n)]
    assert not missing, missing
    assert set(names) == set(_lib.EXPORTED), set(names) ^ set(_lib.EXPORTED)
    assert b"sm_100a" in _lib.lib.spf_version()


def test_header_has_no_torch_types():
    header = open(os.path.join(ROOT, "include", "spurfies_b200.h")).read()
    assert "torch" not in header.split("*/", 1)[1].replace("torch_knnquery", "")
    assert "at::" not in header


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "spurfies_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
