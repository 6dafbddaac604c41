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


def test_header_compiles_as_plain_c_and_cpp(tmp_path):
    """The boundary is a C ABI: the header must be usable from C99 and C++ without CUDA or torch headers."""
    import shutil
    import subprocess
    for cc, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++17", "cpp")):
        if shutil.which(cc) is None:
            continue
        src = tmp_path / f"use_header.{ext}"
        src.write_text('#include "include/spurfies_b200.h"\nint main(void) { const char* (*f)(void) = spf_version; return f == 0; }\n')
        subprocess.check_call([cc, std, "-Wall", "-Werror", "-fsyntax-only", "-I", ROOT, str(src)])
