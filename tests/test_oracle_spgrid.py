"""Pins the oracle's SPGrid addressing (a1/a2) against golden vectors produced by the reference's own SPGrid core,
and, when oracle/_ref is present, against that library directly."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden(fp32):
    return np.load(os.path.join(ROOT, "tests", "golden", "spgrid_fp32.npz" if fp32 else "spgrid_fp64.npz"))


@pytest.mark.parametrize("fp32", [False, True])
def test_mask_constants(oracle, fp32):
    g = _golden(fp32)
    info, masks = oracle.mask_info(fp32)
    assert info == list(g["info"])
    # SURVEY.md A.1 table (obtained from the compiled reference core)
    want = ([0x4924924924924c00, 0x2492492492492300, 0x92492492492490c0] if fp32
            else [0x9249249249249800, 0x4924924924924600, 0x2492492492492180])
    assert masks == want


@pytest.mark.parametrize("fp32", [False, True])
def test_linear_offset_and_back(oracle, fp32):
    g = _golden(fp32)
    off = oracle.linear_offset(g["ijk"], fp32)
    assert (off == g["off"]).all()
    assert (oracle.linear_to_coord(g["off"], fp32) == g["ijk"]).all()


@pytest.mark.parametrize("fp32", [False, True])
def test_packed_add(oracle, fp32):
    g = _golden(fp32)
    assert (oracle.packed_add(g["add_a"], g["add_b"], fp32) == g["add_sum"]).all()


@pytest.mark.parametrize("fp32", [False, True])
def test_page_activation_order(oracle, fp32):
    g = _golden(fp32)
    blocks = oracle.activate(g["group_offsets"], fp32)
    assert len(blocks) == len(g["blocks"])
    assert (blocks == g["blocks"]).all()  # ORDER matters: it defines the DOF numbering


def test_known_offsets(oracle):
    # values printed by the compiled reference
    assert list(oracle.linear_offset([[1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 0, 0]])) == [0x800, 0x200, 0x80, 0x1000]
    assert list(oracle.linear_offset([[1, 0, 0], [0, 1, 0], [0, 0, 1], [2, 0, 0]], True)) == [0x400, 0x100, 0x40, 0x800]
    assert int(oracle.linear_offset([[4095, 4095, 4095]])[0]) == 0x7ffffffff80
    assert int(oracle.linear_offset([[4095, 4095, 4095]], True)[0]) == 0x3ffffffffc0


@pytest.mark.parametrize("fp32", [False, True])
def test_against_compiled_reference(oracle, fp32):
    so = os.path.join(ROOT, "oracle", "_ref", "libspgrid_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference absent); golden vectors cover this")
    ref = C.CDLL(so)
    rng = np.random.default_rng(7)
    ijk = rng.integers(0, 4096, size=(20000, 3), dtype=np.int32)
    want = np.empty(len(ijk), dtype=np.uint64)
    ref.spgrid_ref_linear_offset(int(fp32), C.c_long(len(ijk)), ijk.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
    assert (oracle.linear_offset(ijk, fp32) == want).all()
