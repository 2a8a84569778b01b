"""Pins the restart-file layout (f4: MpmSimulationBase::writeState / readState, Lib/MPM/MpmSimulationBase.cpp:755-785) to the REFERENCE'S OWN code:
tests/golden/restart_ref.npz holds a byte stream written by the reference's DataManager / DataArray / BinaryIO writers and CorotatedIsotropic::write,
compiled where they lie (oracle/restart_ref_shim.cpp -> oracle/_ref/librestart_ref.so; tests/golden/make_restart_golden.py).  The product's
serialisation (writeRestart / readRestart of include/hot_b200_host.hpp, under MpmSimulationB200::writeState / readState; driver tests/cpp/restart_ref.cpp,
no device call) must read that stream back bit-exactly, and must write a stream the reference's own reader (DataManager::readData) reads back bit-exactly and
that holds the same arrays with the same headers (the reference iterates an unordered_map: the ORDER of the arrays is free, its reader looks them up by name)."""
import importlib.util
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_restart_golden", os.path.join(ROOT, "tests", "golden", "make_restart_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
G = np.load(os.path.join(ROOT, "tests", "golden", "restart_ref.npz"))
A = {k: G["in_" + k] for k in gen.KEYS}


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("restart_ref") / "restart_ref")
    lib = os.path.join(ROOT, "hot_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "restart_ref.cpp"),
                           "-o", exe, "-L", lib, "-lhot_b200", f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def _arrays_file(path, a):
    with open(path, "wb") as f:
        f.write(struct.pack("<q", len(a["m"])))
        for k in gen.KEYS:
            f.write(np.ascontiguousarray(a[k], dtype=np.float64).tobytes())


def _read_arrays_file(path):
    raw = open(path, "rb").read()
    n = struct.unpack_from("<q", raw, 0)[0]
    out, pos = {}, 8
    for k in gen.KEYS:
        w = gen.WIDTH[k]
        v = np.frombuffer(raw, dtype=np.float64, count=n * w, offset=pos); pos += 8 * n * w
        out[k] = v.reshape(n, w) if w > 1 else v
    return n, out


def _parse(b):
    """{array name: (lg2_grain_size, ranges, size, element bytes, payload)} + the bytes after the arrays, following DataManager.h:263-273"""
    b = bytes(b)
    pos = 0

    def take(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, b, pos); pos += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v
    n, na = take("i"), take("Q")
    arrays = {}
    for _ in range(na):
        ln = take("Q"); name = b[pos:pos + ln].decode(); pos += ln
        grain = take("i")
        nr, rb = take("Q"), take("Q")
        ranges = struct.unpack_from(f"<{2 * nr}i", b, pos); pos += nr * rb
        size, eb = take("Q"), take("Q")
        payload = b[pos:pos + size * eb]; pos += size * eb
        if name == "CorotatedIsotropic":            # 24 raw bytes of {bool project; double mu, lambda} per entry: the 7 padding bytes are indeterminate
            payload = b"".join(payload[24 * i:24 * i + 1] + payload[24 * i + 8:24 * i + 24] for i in range(size))
        arrays[name] = (grain, ranges, size, eb, payload)
    return n, arrays, b[pos:]


def test_product_reader_reads_the_reference_stream(driver, tmp_path):
    src, dst = str(tmp_path / "restart.dat"), str(tmp_path / "arrays.bin")
    open(src, "wb").write(bytes(G["bytes"]))
    subprocess.check_call([driver, "read", src, dst])
    n, back = _read_arrays_file(dst)
    assert n == len(A["m"])
    for k in gen.KEYS:
        assert np.array_equal(back[k], A[k]), k


def test_product_writer_writes_the_reference_layout(driver, tmp_path):
    src, dst = str(tmp_path / "arrays.bin"), str(tmp_path / "restart.dat")
    _arrays_file(src, A)
    subprocess.check_call([driver, "write", src, dst])
    ours = open(dst, "rb").read()
    n_o, arr_o, tail_o = _parse(ours)
    n_r, arr_r, tail_r = _parse(G["bytes"])
    assert n_o == n_r == len(A["m"]) and len(ours) == len(G["bytes"])
    assert arr_o == arr_r                       # every array: grain size, ranges, size, element bytes and payload identical (the order is free)
    assert tail_o == tail_r                     # the two empty mesh index vectors


@pytest.mark.skipif(not os.path.exists(gen.REF_LIB), reason="oracle/_ref/librestart_ref.so not built (needs /root/reference)")
def test_reference_reader_reads_the_product_stream(driver, tmp_path):
    src, dst = str(tmp_path / "arrays.bin"), str(tmp_path / "restart.dat")
    _arrays_file(src, A)
    subprocess.check_call([driver, "write", src, dst])
    n, back = gen.reference_read(np.frombuffer(open(dst, "rb").read(), dtype=np.uint8))
    assert n == len(A["m"])
    for k in gen.KEYS:
        assert np.array_equal(back[k], A[k]), k
    assert _parse(gen.reference_write(A)) == _parse(G["bytes"])          # (the reference's writer reproduces the golden stream up to the padding bytes)
