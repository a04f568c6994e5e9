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
