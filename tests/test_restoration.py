"""Loop restoration (self-guided filter) of the oracle, pinned against libaom 3.13.1's C functions
av1_selfguided_restoration_c / av1_apply_selfguided_restoration_c (normative arithmetic, AV1 spec 7.17.3-7.17.4)."""
import ctypes
import numpy as np
import pytest
import aomlib
import harness as H

P32 = ctypes.POINTER(ctypes.c_int32)


def _ext_plane(rng, w, h, bd, smooth):
    """(h+6) x (w+6) samples: the region plus the 3-sample border the filters read"""
    if smooth:
        yy, xx = np.mgrid[0:h + 6, 0:w + 6]
        v = 0.5 + 0.35 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + rng.normal(0, 0.02, (h + 6, w + 6))
    else:
        v = rng.random((h + 6, w + 6))
    return np.clip(v * ((1 << bd) - 1), 0, (1 << bd) - 1).astype(np.uint16)


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("set_idx", [0, 2, 6, 9, 11, 13, 14, 15])
def test_box_filters_match_libaom(bd, set_idx):
    O = H.oracle()
    fn = aomlib.func("av1_selfguided_restoration_c", ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                     P32, P32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int)
    rng = np.random.default_rng(bd * 100 + set_idx)
    for (w, h, smooth) in ((64, 64, True), (33, 18, False), (8, 56, True)):
        ext = _ext_plane(rng, w, h, bd, smooth)
        stride = w + 6
        a0 = np.zeros((h, w), np.int32); a1 = np.zeros((h, w), np.int32)
        if bd == 8:
            buf = ext.astype(np.uint8)
            ptr = buf.ctypes.data + 3 * stride + 3
        else:
            buf = ext
            ptr = (buf.ctypes.data + 2 * (3 * stride + 3)) >> 1          # CONVERT_TO_BYTEPTR
        fn(ptr, w, h, stride, a0.ctypes.data_as(P32), a1.ctypes.data_as(P32), w, set_idx, bd, int(bd > 8))
        o0 = np.zeros((h, w), np.int32); o1 = np.zeros((h, w), np.int32)
        O.orc_sgr_region_plane(ext.ctypes.data, ext.ctypes.data, stride, w + 6, h + 6, bd, -64, 1 << 20, 3, 3, w, h, set_idx,
                               o0.ctypes.data, o1.ctypes.data)
        r0, r1 = H.SGR_PARAMS[set_idx][0], H.SGR_PARAMS[set_idx][1]
        if r0:
            assert np.array_equal(a0, o0), "r=2 pass, set %d" % set_idx
        if r1:
            assert np.array_equal(a1, o1), "r=1 pass, set %d" % set_idx


@pytest.mark.parametrize("bd", [8, 10])
def test_projection_matches_libaom(bd):
    O = H.oracle()
    fn = aomlib.func("av1_apply_selfguided_restoration_c", ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                     ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_int, P32, ctypes.c_int, ctypes.c_int)
    rng = np.random.default_rng(bd)
    w, h = 48, 40
    tmp = np.zeros(1 << 20, np.int32)
    for set_idx, xqd in ((2, (-30, 60)), (6, (10, 40)), (11, (0, 70)), (14, (-50, 95)), (9, (-96, -32)), (0, (31, 95))):
        ext = _ext_plane(rng, w, h, bd, True)
        stride = w + 6
        x = (ctypes.c_int * 2)(*xqd)
        if bd == 8:
            buf = ext.astype(np.uint8); dst = np.zeros((h, w), np.uint8)
            fn(buf.ctypes.data + 3 * stride + 3, w, h, stride, set_idx, x, dst.ctypes.data, w, tmp.ctypes.data_as(P32), bd, 0)
        else:
            buf = ext; dst = np.zeros((h, w), np.uint16)
            fn((buf.ctypes.data + 2 * (3 * stride + 3)) >> 1, w, h, stride, set_idx, x, dst.ctypes.data >> 1, w, tmp.ctypes.data_as(P32), bd, 1)
        out = np.zeros((h, w), np.uint16)
        O.orc_sgr_apply_region(ext.ctypes.data, stride, w + 6, h + 6, bd, 3, 3, w, h, set_idx, xqd[0], xqd[1], out.ctypes.data)
        assert np.array_equal(dst.astype(np.uint16), out), "set %d xqd %s" % (set_idx, xqd)


def test_unit_geometry():
    O = H.oracle()
    for size, n in ((1, 1), (31, 1), (64, 1), (95, 1), (96, 2), (128, 2), (159, 2), (160, 3), (1080, 17), (1920, 30)):
        assert max((size + 32) // 64, 1) == n


# ---- whole pipeline: oracle encode -> deblock -> CDEF -> loop restoration search/apply -> pack -> dav1d ---------------
import dav1d


@pytest.mark.parametrize("w,h,depth,planes,q,tiles", [
    (64, 64, 8, 3, 121, 1),          # one unit
    (200, 136, 8, 3, 121, 1),        # 3 x 2 units, merged last row (136 -> 2 unit rows, the second 72 + 8 rows tall)
    (161, 100, 10, 3, 147, 2),       # odd sizes: clamping at the right / bottom frame edge, two tiles (reference reset)
    (95, 95, 8, 1, 100, 1),          # a single unit larger than 64 x 64 (two stripes), monochrome
    (384, 256, 10, 3, 66, 4),
    (130, 250, 8, 3, 199, 1),
])
def test_restoration_stream_decodes_sample_exact(w, h, depth, planes, q, tiles):
    if not dav1d.available():
        pytest.skip("dav1d (Pillow libavif) not found")
    f = H.frame_setup(w, h, depth, planes, q, tiles)
    f.enable_restoration = 1
    for p in range(planes):
        f.lr_type[p] = 1
    rt = H.rate_tables(f.base_q_idx)
    img = H.synth_image(w, h, w * 3 + h)
    src = H.pad_planes([img[..., c].astype(np.uint16) * (4 if depth == 10 else 1) for c in range(planes)], f)
    rec = H.oracle_encode(f, rt, src)
    fin, deb = H.oracle_filters(f, rec, src)
    used = sum(int(rec.lr[p]["use"].sum()) for p in range(planes))
    assert used > 0, "the search never switched restoration on: the test would prove nothing"
    obu = H.pack(f, rec)
    dec, bpc, layout = dav1d.decode(obu, 7)
    assert bpc == depth
    for p in range(planes):
        assert np.array_equal(dec[p], fin[p][:h, :w]), "plane %d after deblock + CDEF + LR" % p
    # the restoration must not make things worse than the CDEF output it started from
    for p in range(planes):
        e_cdef = float(((rec.cdef_out[p][:h, :w].astype(np.int64) - src[p][:h, :w]) ** 2).sum())
        e_lr = float(((fin[p][:h, :w].astype(np.int64) - src[p][:h, :w]) ** 2).sum())
        assert e_lr <= e_cdef * 1.001
