"""Decoder conformance: oracle hot path -> product symbol packer -> dav1d 1.5.3.
The decoded planes must equal the encoder's own reconstruction SAMPLE-EXACTLY; this pins predictors,
dequantisation, inverse transforms, reconstruction, every CDF/context and the range coder."""
import numpy as np
import pytest
import dav1d
import harness as H

pytestmark = pytest.mark.skipif(not dav1d.available(), reason="dav1d not available")


def roundtrip(w, h, depth=8, planes=3, q=121, tiles=1, ymask=0x1fff, uvmask=0x1fff, txmask=0x1f, pmin=4, pmax=16, rdo_tx=1,
              seed=0, reduced=1, ief=0):
    f = H.frame_setup(w, h, depth, planes, q, tiles)
    f.y_mode_mask = ymask; f.uv_mode_mask = uvmask; f.tx_type_mask = txmask
    f.partition_min = pmin; f.partition_max = pmax; f.rdo_tx_decision = rdo_tx; f.reduced_tx_set = reduced
    f.enable_intra_edge_filter = ief
    rt = H.rate_tables(f.base_q_idx)
    img = H.synth_image(w, h, seed)
    pl = [img[..., c].astype(np.uint16) * (4 if depth == 10 else 1) for c in range(planes)]
    src = H.pad_planes(pl, f)
    rec = H.oracle_encode(f, rt, src)
    final, deblocked = H.oracle_filters(f, rec, src)
    obu = H.pack(f, rec)
    # three decodes: in-loop filters off / deblocking only / everything -> pins each stage separately
    for flt, ref, what in ((0, rec.rec, "reconstruction"), (dav1d.INLOOP_DEBLOCK, deblocked, "deblocked"), (dav1d.INLOOP_ALL, final, "CDEF output")):
        dec, bpc, layout = dav1d.decode(obu, flt)
        assert bpc == depth and layout == (3 if planes == 3 else 0)
        for p in range(planes):
            assert np.array_equal(dec[p], ref[p][:h, :w]), "plane %d: dav1d differs from the encoder's %s" % (p, what)
    mse = np.mean((final[0][:h, :w].astype(float) - pl[0]) ** 2)
    return len(obu), 10 * np.log10(((1 << depth) - 1) ** 2 / max(mse, 1e-9))


@pytest.mark.parametrize("mode", range(13))
def test_each_luma_predictor(mode):
    roundtrip(136, 104, ymask=1 << mode, uvmask=1, pmin=16, rdo_tx=0, txmask=1)


@pytest.mark.parametrize("mode", range(13))
def test_each_chroma_predictor_and_derived_tx_type(mode):
    roundtrip(72, 56, ymask=1, uvmask=1 << mode, pmin=8, pmax=8, rdo_tx=0, txmask=1)


@pytest.mark.parametrize("size", [4, 8, 16, 32])
@pytest.mark.parametrize("tx", range(5))
def test_each_transform_type_and_size(size, tx):
    if size == 32 and tx != 0:
        pytest.skip("32x32 intra blocks are DCT only")
    roundtrip(96, 64, txmask=1 << tx, pmin=size, pmax=size, rdo_tx=0)


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("size", [4, 8])
@pytest.mark.parametrize("tx", [5, 6])
def test_one_dimensional_dct_types_of_the_7_type_set(size, tx, depth):
    """V_DCT / H_DCT (intra set of transforms up to 8x8 when reduced_tx_set is off, av1encoder.rs:581): their own scans (mrow /
    mcol), coefficient contexts of the 1-D transform classes and the 7-symbol tx-type CDF, all decoded by dav1d"""
    roundtrip(96, 64, depth=depth, txmask=1 << tx, pmin=size, pmax=size, rdo_tx=0, reduced=0)


def test_seven_type_search_beats_or_equals_five_types():
    """full search with the 7-type set on 4x4 / 8x8 transforms: decodes, and the RD search never does worse for having more types"""
    s7, p7 = roundtrip(200, 136, txmask=0x7f, pmin=4, pmax=8, reduced=0, q=100)
    s5, p5 = roundtrip(200, 136, txmask=0x1f, pmin=4, pmax=8, reduced=0, q=100)
    assert s7 <= s5 * 1.01 or p7 >= p5


@pytest.mark.parametrize("w,h", [(8, 8), (16, 300), (1000, 24), (129, 101), (128, 85), (72, 40), (200, 136)])
def test_frame_edges_and_forced_splits(w, h):
    roundtrip(w, h)


@pytest.mark.parametrize("depth,planes", [(8, 3), (10, 3), (8, 1), (10, 1)])
def test_full_search_depth_and_monochrome(depth, planes):
    size, psnr = roundtrip(300, 200, depth=depth, planes=planes, tiles=4)
    assert psnr > 30


def test_many_tiles_and_quantisers():
    roundtrip(640, 360, depth=10, tiles=12)
    roundtrip(256, 256, q=200, pmax=32)
    roundtrip(256, 192, q=20)
    roundtrip(256, 192, q=255)
    roundtrip(256, 192, q=1, pmin=8)


def test_full_tx_set_signalling():
    roundtrip(136, 104, reduced=0)


@pytest.mark.parametrize("lf,cdef", [((20, 25, 15, 30), (15 * 4 + 3, 9 * 4 + 2, 5)), ((63, 40, 50, 10), (7 * 4 + 1, 3, 6)), ((0, 0, 0, 0), (2, 12, 3))])
def test_forced_filter_strengths(lf, cdef):
    """deblock levels / CDEF strengths far from the closed-form picks, 8 and 10 bit"""
    for depth in (8, 10):
        f = H.frame_setup(200, 136, depth, 3, 150, 2)
        for i in range(4):
            f.lf_level[i] = lf[i]
        f.cdef_bits = 0; f.cdef_y_strength[0] = cdef[0]; f.cdef_uv_strength[0] = cdef[1]; f.cdef_damping = cdef[2]
        img = H.synth_image(200, 136, 11)
        src = H.pad_planes([img[..., c].astype(np.uint16) * (4 if depth == 10 else 1) for c in range(3)], f)
        rec = H.oracle_encode(f, H.rate_tables(f.base_q_idx), src)
        final, deb = H.oracle_filters(f, rec, src)
        obu = H.pack(f, rec)
        for flt, ref in ((dav1d.INLOOP_DEBLOCK, deb), (dav1d.INLOOP_ALL, final)):
            dec, _, _ = dav1d.decode(obu, flt)
            for p in range(3):
                assert np.array_equal(dec[p], ref[p][:136, :200])


def test_rate_increases_with_quality():
    s_lo, p_lo = roundtrip(256, 192, q=180)
    s_hi, p_hi = roundtrip(256, 192, q=60)
    assert s_hi > s_lo and p_hi > p_lo


@pytest.mark.parametrize("w,h,depth,planes,q,tiles", [
    (256, 256, 8, 3, 180, 1), (200, 136, 10, 3, 150, 2), (95, 95, 8, 1, 100, 1), (384, 320, 8, 3, 220, 4), (130, 70, 10, 3, 199, 1),
])
def test_32x32_partitions(w, h, depth, planes, q, tiles):
    """partition range 4..32 (ravif speed <= 2): BLOCK_32X32 / TX_32X32 (DCT only), forced splits at the frame edges"""
    n, psnr = roundtrip(w, h, depth=depth, planes=planes, q=q, tiles=tiles, pmax=32, seed=w)
    assert n > 0
    n8, _ = roundtrip(w, h, depth=depth, planes=planes, q=q, tiles=tiles, pmin=8, pmax=32, seed=w)
    n16, _ = roundtrip(w, h, depth=depth, planes=planes, q=q, tiles=tiles, pmin=16, pmax=32, rdo_tx=0, seed=w)
    assert n8 > 0 and n16 > 0


@pytest.mark.parametrize("mode", range(1, 9))
@pytest.mark.parametrize("size", [4, 8, 16, 32])
def test_intra_edge_filter_each_directional_mode_and_size(mode, size):
    """enable_intra_edge_filter = 1: every directional mode (with its refined angle deltas) at every block size, so that
    every strength class, the corner filter (>= 16x16) and the edge upsampling (<= 8x8) are decoded by dav1d"""
    pmin = min(size, 8) if size >= 8 else 4
    for depth, q in ((8, 100), (10, 160)):
        n, _ = roundtrip(136, 104, depth=depth, q=q, ymask=(1 << mode) | 1, uvmask=0x1fff, pmin=pmin if size < 16 else size, pmax=size,
                         rdo_tx=0, seed=mode * 10 + size, ief=1)
        assert n > 0


@pytest.mark.parametrize("w,h,depth,planes,q,tiles,pmax", [(200, 136, 8, 3, 121, 2, 16), (161, 100, 10, 3, 147, 1, 32), (95, 95, 8, 1, 100, 1, 16),
                                                        (256, 192, 8, 3, 60, 4, 32)])
def test_intra_edge_filter_full_search(w, h, depth, planes, q, tiles, pmax):
    """all tools on (smooth neighbours change the filter type of the closed-loop reconstruction)"""
    n, _ = roundtrip(w, h, depth=depth, planes=planes, q=q, tiles=tiles, pmax=pmax, seed=w, ief=1)
    assert n > 0
