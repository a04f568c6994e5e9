"""AVIF container written by avif_mux.cpp: ftypavif at offset 4 (reference tests/stdio.rs:23), parsed and
decoded by Pillow's libavif; decoded picture == dav1d's decode of the same payload."""
import ctypes, io
import numpy as np
import pytest
from PIL import Image
import dav1d
import harness as H


class MuxIn(ctypes.Structure):
    _fields_ = [("color", ctypes.c_void_p), ("color_len", ctypes.c_size_t), ("alpha", ctypes.c_void_p), ("alpha_len", ctypes.c_size_t),
                ("exif", ctypes.c_void_p), ("exif_len", ctypes.c_size_t), ("width", ctypes.c_uint32), ("height", ctypes.c_uint32),
                ("depth", ctypes.c_int), ("matrix_coefficients", ctypes.c_int), ("premultiplied_alpha", ctypes.c_int),
                ("color_level_idx", ctypes.c_int), ("alpha_level_idx", ctypes.c_int)]


def encode_obu(planes, w, h, depth, q=100):
    f = H.frame_setup(w, h, depth, len(planes), q, 1)
    if len(planes) == 3:
        f.color_description_present = 1; f.color_primaries = 1; f.transfer_characteristics = 13; f.matrix_coefficients = 0
    src = H.pad_planes(planes, f)
    rec = H.oracle_encode(f, H.rate_tables(f.base_q_idx), src)
    rec.final, _ = H.oracle_filters(f, rec, src)
    return H.pack(f, rec), rec, f


@pytest.mark.parametrize("with_alpha", [False, True])
def test_mux_is_parsed_by_libavif(with_alpha):
    w, h = 96, 64
    img = H.synth_image(w, h, 5, alpha=with_alpha)
    # identity matrix (GBR planes) so that the decoded RGB can be compared exactly with the reconstruction
    col, rec, f = encode_obu([img[..., 1], img[..., 2], img[..., 0]], w, h, 8, q=40)
    mi = MuxIn()
    cb = ctypes.create_string_buffer(col, len(col))
    mi.color = ctypes.addressof(cb); mi.color_len = len(col)
    if with_alpha:
        al, arec, af = encode_obu([img[..., 3]], w, h, 8, q=40)
        ab = ctypes.create_string_buffer(al, len(al))
        mi.alpha = ctypes.addressof(ab); mi.alpha_len = len(al)
    mi.width = w; mi.height = h; mi.depth = 8; mi.matrix_coefficients = 0
    L = H.hostlib()
    mi.color_level_idx = L.b200av1_seq_level_idx(ctypes.byref(f)); mi.alpha_level_idx = mi.color_level_idx
    out = ctypes.c_void_p(); n = ctypes.c_size_t()
    assert L.b200avif_mux(ctypes.byref(mi), ctypes.byref(out), ctypes.byref(n)) == 0
    data = ctypes.string_at(out.value, n.value)
    assert data[4:12] == b"ftypavif"
    im = Image.open(io.BytesIO(data))
    assert im.size == (w, h)
    assert im.mode == ("RGBA" if with_alpha else "RGB")
    arr = np.asarray(im.convert("RGBA" if with_alpha else "RGB"))
    g, b, r = [rec.final[p][:h, :w] for p in range(3)]
    assert np.array_equal(arr[..., 0], r) and np.array_equal(arr[..., 1], g) and np.array_equal(arr[..., 2], b)
    if with_alpha:
        assert np.array_equal(arr[..., 3], arec.final[0][:h, :w])


def test_exif_item_round_trips_through_libavif():
    """Encoder::with_exif / set_exif (av1encoder.rs:206-219) -> Aviffy::set_exif (:470-472): the Exif item written by the
    muxer is found and returned byte-for-byte by libavif (Pillow), and the picture still decodes."""
    w, h = 64, 48
    img = H.synth_image(w, h, 11)
    col, rec, f = encode_obu([img[..., 1], img[..., 2], img[..., 0]], w, h, 8, q=60)
    # a minimal TIFF-structured Exif payload: little-endian header, one IFD0 entry (Software = "b20", inline ASCII);
    # (an Orientation tag would be rewritten by Pillow, which takes the orientation from irot / imir)
    exif = b"II*\x00\x08\x00\x00\x00" + b"\x01\x00" + b"\x31\x01\x02\x00\x04\x00\x00\x00b20\x00" + b"\x00\x00\x00\x00"
    mi = MuxIn()
    cb = ctypes.create_string_buffer(col, len(col)); eb = ctypes.create_string_buffer(exif, len(exif))
    mi.color = ctypes.addressof(cb); mi.color_len = len(col)
    mi.exif = ctypes.addressof(eb); mi.exif_len = len(exif)
    mi.width = w; mi.height = h; mi.depth = 8; mi.matrix_coefficients = 0
    L = H.hostlib()
    mi.color_level_idx = L.b200av1_seq_level_idx(ctypes.byref(f)); mi.alpha_level_idx = mi.color_level_idx
    out = ctypes.c_void_p(); n = ctypes.c_size_t()
    assert L.b200avif_mux(ctypes.byref(mi), ctypes.byref(out), ctypes.byref(n)) == 0
    data = ctypes.string_at(out.value, n.value)
    assert b"Exif" in data and exif in data
    im = Image.open(io.BytesIO(data))
    assert im.size == (w, h)
    got = im.info.get("exif")
    assert got is not None and bytes(got).endswith(exif), "libavif did not return the Exif item"
    assert im.getexif().get(0x0131) == "b20"
    arr = np.asarray(im.convert("RGB"))
    # Pillow applies no orientation by itself: the decoded samples are the reconstruction
    assert np.array_equal(arr[..., 1], rec.final[0][:h, :w])
