"""CPU: the C-ABI library loads and exports every symbol include/fpose.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fpose.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from foundationpose_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in fpose.h but not exported: {missing}"


def test_error_string_and_counters_without_gpu():
    from foundationpose_b200 import _lib

    assert _lib.launch_count() >= 0
    assert isinstance(_lib.lib.fp_last_error(), (bytes, type(None)))


def test_product_path_has_no_cpu_fallback():
    """The engine refuses to run without a CUDA device instead of silently falling back."""
    import pytest
    import torch

    from foundationpose_b200 import _lib
    from foundationpose_b200.engine import Engine

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.FposeError):
        Engine()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "foundationpose_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
