"""The C-ABI library loads on a CPU-only box and exports every symbol include/hot_b200.h declares
(no compute calls without a GPU); the product path fails loudly without a device."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from hot_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 20
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/hot_b200.h but not exported: {missing}"


def test_binding_covers_header():
    from hot_b200 import _lib
    lib = _lib.load_library()
    assert sorted(lib._hot_signatures) == _lib.declared_symbols()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import hot_b200
    with pytest.raises(hot_b200.HotError):
        hot_b200.MpmSimulationB200(dx=0.1)


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hot_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libhot_oracle" not in text and "oracle_binding" not in text and "orc_" not in text, f


def test_host_mirror_header_compiles():
    """include/hot_b200_host.hpp (the C++ mirror of the reference's operator surface) must stay a valid, warning-free C++17
    translation unit against include/hot_b200.h; the GPU test links and runs it."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["g++", "-std=c++17", "-O0", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-fsyntax-only",
                        os.path.join(root, "tests", "cpp", "host_step.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
