"""Generates tests/golden/spgrid_{fp64,fp32}.npz from the REFERENCE's own SPGrid core
(oracle/_ref/libspgrid_ref.so, compiled by oracle/Makefile from /root/reference/Lib/SPGrid/Core where it lies).
Run in the authoring container only (the reference does not exist on the GPU box):
    make -C oracle ref && python tests/golden/make_spgrid_golden.py
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libspgrid_ref.so"))
ref.spgrid_ref_activate.restype = C.c_long
p = lambda a: a.ctypes.data_as(C.c_void_p)


def main():
    rng = np.random.default_rng(20260117)
    for fp32 in (0, 1):
        info = (C.c_int * 6)()
        ref.spgrid_ref_info(fp32, info)
        info = np.array(list(info), dtype=np.int32)
        # coordinates: corners, powers of two +-1, and random points of the 4096^3 box
        special = [0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 31, 32, 63, 64, 127, 128, 255, 256, 511, 512, 1023, 1024, 2047, 2048, 4094, 4095]
        grid = np.array([(a, b, c) for a in special for b in (0, 1, 5, 4095) for c in (0, 3, 2048, 4095)], dtype=np.int32)
        ijk = np.concatenate([grid, rng.integers(0, 4096, size=(4000, 3), dtype=np.int32)]).astype(np.int32)
        off = np.empty(len(ijk), dtype=np.uint64)
        ref.spgrid_ref_linear_offset(fp32, C.c_long(len(ijk)), p(ijk), p(off))
        back = np.empty_like(ijk)
        ref.spgrid_ref_linear_to_coord(fp32, C.c_long(len(ijk)), p(off), p(back))
        assert (back == ijk).all()
        # packed add of stencil offsets (0..2)^3 and block strides onto random bases away from the upper wall
        base_ijk = rng.integers(0, 4090, size=(2000, 3), dtype=np.int32)
        d_ijk = rng.integers(0, 5, size=(2000, 3), dtype=np.int32)
        a = np.empty(2000, dtype=np.uint64); b = np.empty(2000, dtype=np.uint64); s = np.empty(2000, dtype=np.uint64)
        ref.spgrid_ref_linear_offset(fp32, C.c_long(2000), p(base_ijk), p(a))
        ref.spgrid_ref_linear_offset(fp32, C.c_long(2000), p(d_ijk), p(b))
        ref.spgrid_ref_packed_add(fp32, C.c_long(2000), p(a), p(b), p(s))
        # page activation replayed through the real SPGrid_Page_Map: a blob of group offsets in sorted order
        cells = rng.integers(40, 90, size=(600, 3), dtype=np.int32)
        goff = np.empty(600, dtype=np.uint64)
        ref.spgrid_ref_linear_offset(fp32, C.c_long(600), p(cells), p(goff))
        goff = np.unique((goff >> np.uint64(12)) << np.uint64(12))  # one offset per page, ascending = sorted key order
        blocks = np.empty(9 * len(goff) + 8, dtype=np.uint64)
        nb = ref.spgrid_ref_activate(fp32, C.c_long(len(goff)), p(goff), p(blocks), C.c_long(len(blocks)))
        out = os.path.join(ROOT, "tests", "golden", "spgrid_fp32.npz" if fp32 else "spgrid_fp64.npz")
        np.savez_compressed(out, info=info, ijk=ijk, off=off, add_a=a, add_b=b, add_sum=s, group_offsets=goff,
                            blocks=blocks[:nb].copy())
        print(out, "info", info, "pages", nb)


if __name__ == "__main__":
    main()
