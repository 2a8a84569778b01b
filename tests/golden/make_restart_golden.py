"""Generates tests/golden/restart_ref.npz (f4, restart files): a restart byte stream written by the REFERENCE'S OWN DataManager / DataArray / BinaryIO code
and CorotatedIsotropic::write (Lib/Ziran/CS/DataStructure, Lib/Ziran/CS/Util/BinaryIO.h), compiled where they lie (oracle/restart_ref_shim.cpp ->
oracle/_ref/librestart_ref.so) and driven the way MpmSimulationBase::writeState / Scene::writeState drive them for an MPM scene, for a seeded particle set.
tests/test_restart_ref.py reads it with the product's reader and compares the product's writer with it.
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_restart_golden.py"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "librestart_ref.so")
OUT = os.path.join(ROOT, "tests", "golden", "restart_ref.npz")
KEYS = ("X", "V", "m", "vol", "F", "mu", "lam")
WIDTH = dict(X=3, V=3, m=1, vol=1, F=9, mu=1, lam=1)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def particles(n=57, seed=3):
    rng = np.random.default_rng(seed)
    return dict(X=rng.random((n, 3)), V=rng.random((n, 3)) - 0.5, m=1e-3 * (1 + rng.random(n)), vol=1e-6 * (1 + rng.random(n)),
                F=np.eye(3).reshape(1, 9) + 0.1 * (rng.random((n, 9)) - 0.5), mu=1e4 * (1 + rng.random(n)), lam=2e4 * (1 + rng.random(n)))


def reference_write(a):
    lib = C.CDLL(REF_LIB)
    lib.zr_restart_write.restype = C.c_long
    n = len(a["m"])
    arrs = [np.ascontiguousarray(a[k], dtype=np.float64) for k in KEYS]
    buf = np.zeros(1 << 20, dtype=np.uint8)
    ln = lib.zr_restart_write(C.c_long(n), *[_p(x) for x in arrs], _p(buf), C.c_long(len(buf)))
    assert ln > 0
    return buf[:ln].copy()


def reference_read(b, cap=4096):
    lib = C.CDLL(REF_LIB)
    lib.zr_restart_read.restype = C.c_long
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = {k: np.empty((cap, WIDTH[k])) if WIDTH[k] > 1 else np.empty(cap) for k in KEYS}
    n = lib.zr_restart_read(_p(b), C.c_long(len(b)), C.c_long(cap), *[_p(out[k]) for k in KEYS])
    return n, {k: v[:max(n, 0)] for k, v in out.items()}


if __name__ == "__main__":
    a = particles()
    b = reference_write(a)
    n, back = reference_read(b)
    assert n == len(a["m"]) and all(np.array_equal(back[k], a[k]) for k in KEYS)
    np.savez_compressed(OUT, bytes=b, **{"in_" + k: a[k] for k in KEYS})
    print("wrote", OUT, "restart stream of", len(b), "bytes for", n, "particles")
