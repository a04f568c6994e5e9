"""Intra edge filter / upsampling helpers of the oracle pinned against libaom 3.13.1 (av1_highbd_filter_intra_edge_c,
av1_highbd_upsample_intra_edge_c and the static strength / upsample selectors re-derived from the AV1 spec tables)."""
import ctypes
import numpy as np
import pytest
import aomlib
import harness as H

PU16 = ctypes.POINTER(ctypes.c_uint16)


def _orc():
    L = H.oracle()
    L.orc_ief_filter_hook.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.orc_ief_upsample_hook.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return L


@pytest.mark.parametrize("bd", [8, 10])
def test_edge_filter_matches_libaom(bd):
    L = _orc()
    fn = aomlib.func("av1_highbd_filter_intra_edge_c", None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int)
    rng = np.random.default_rng(bd)
    for sz in (5, 9, 17, 33, 65):
        for strength in (1, 2, 3):
            edge = rng.integers(0, 1 << bd, sz + 8, dtype=np.uint16)
            a = edge.copy(); b = edge.copy()
            fn(a.ctypes.data + 2 * 4, sz, strength)                 # libaom: p[0] = corner, p[1 .. sz-1] filtered
            L.orc_ief_filter_hook(b.ctypes.data, 4, sz, strength)   # oracle: element [off] = corner
            assert np.array_equal(a, b), (sz, strength)


@pytest.mark.parametrize("bd", [8, 10])
def test_edge_upsample_matches_libaom(bd):
    L = _orc()
    fn = aomlib.func("av1_highbd_upsample_intra_edge_c", None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int)
    rng = np.random.default_rng(bd + 7)
    for num_px in (4, 8, 12, 16):
        edge = rng.integers(0, 1 << bd, 64, dtype=np.uint16)
        a = edge.copy(); b = edge.copy()
        off = 8                                                    # element [off] = corner (index -1)
        fn(a.ctypes.data + 2 * (off + 1), num_px, bd)              # libaom: p points at index 0, reads p[-1], writes p[-2 ..]
        L.orc_ief_upsample_hook(b.ctypes.data, off, num_px, bd)
        assert np.array_equal(a, b), num_px


def test_strength_and_upsample_selection_tables():
    """AV1 spec 7.11.2.9 / 7.11.2.10 spot values"""
    L = _orc()
    S = L.orc_ief_strength_hook; U = L.orc_ief_use_upsample_hook
    assert [S(8, 0, d) for d in (0, 55, 56, 90)] == [0, 0, 1, 1]
    assert [S(16, 0, d) for d in (39, 40)] == [0, 1]
    assert [S(24, 0, d) for d in (7, 8, 16, 32)] == [0, 1, 2, 3]
    assert [S(32, 0, d) for d in (0, 1, 4, 32)] == [0, 1, 2, 3]
    assert [S(64, 0, d) for d in (0, 1)] == [0, 3]
    assert [S(8, 1, d) for d in (39, 40, 64)] == [0, 1, 2]
    assert [S(16, 1, d) for d in (19, 20, 48)] == [0, 1, 2]
    assert [S(24, 1, d) for d in (3, 4)] == [0, 3] and [S(40, 1, d) for d in (0, 1)] == [0, 3]
    assert [U(16, 0, d) for d in (0, 1, 39, 40)] == [0, 1, 1, 0] and U(24, 0, 10) == 0
    assert U(8, 1, 10) == 1 and U(16, 1, 10) == 0
