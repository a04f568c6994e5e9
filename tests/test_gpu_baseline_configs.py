"""GPU parity at the geometry and settings of BASELINE.json configs #3, #4, #5 (config #2 lives in
test_gpu_parity.py::test_rdo_full_hd_frame_parity), the narrowest seam b200av1_encode_planes
(encode_to_av1, av1encoder.rs:749-771) and the EXIF item (av1encoder.rs:470-472) end to end.
Every comparison is bit-exact against the CPU oracle / CPU pipeline."""
import ctypes, io, os, sys
import numpy as np
import pytest
from PIL import Image
import dav1d
import harness as H
from test_gpu_parity import device_rdo, check_analysis, check_records, check_filters
from test_gpu_e2e import payloads

sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
pytestmark = pytest.mark.gpu


def C():
    import cavif_rs_b200 as m
    assert m.device_count() > 0
    return m


def test_config3_4k_10bit_frame_parity():
    """config #3: one 3840x2160 frame, q80 (quantizer 121), speed 4, 10-bit, 126 tiles (threads >= cap)."""
    import cpu_encode as CE
    w, h = 3840, 2160
    img = H.synth_image(w, h, 3)
    planes, _, _ = CE.convert(img, 10)
    f = CE.configure(w, h, 10, 3, 121, 4, threads=128)
    assert f.tile_cols * f.tile_rows >= 126 and f.bit_depth == 10 and f.partition_max == 16
    src = H.pad_planes(planes, f)
    ref = H.oracle_encode(f, H.rate_tables(f.base_q_idx), src)
    ref_fin, ref_deb = H.oracle_filters(f, ref, src)
    out = device_rdo(f, src)
    check_analysis(ref, out, f)
    check_records(ref, out, 3)
    check_filters(f, ref, ref_fin, ref_deb, out, 3, w, h)
    assert H.pack(f, out) == H.pack(f, ref)


def test_config4_alpha_batch_through_chunked_pipeline():
    """config #4: 1920x1080 RGBA with real alpha, q60 speed 6, depth Auto (10-bit), alpha quality 80, UnassociatedClean,
    through encode_batch of 16 images (4 chunks on 3 contexts): sampled images byte-identical to the CPU pipeline and to
    the same image encoded alone."""
    import cpu_encode as CE
    m = C()
    w, h = 1920, 1080
    base = [H.synth_image(w, h, 40 + i, alpha=True) for i in range(4)]
    imgs = [np.ascontiguousarray(np.roll(base[i % 4], (53 * (i // 4), 29 * (i // 4)), axis=(0, 1))) for i in range(16)]
    enc = m.Encoder.new().with_quality(60).with_speed(6).with_num_threads(128)
    res = enc.encode_batch(imgs)
    assert len(res) == 16
    for i in (0, 5, 15):            # first chunk, a middle chunk, the last image of the last chunk
        ref = CE.encode_rgba(imgs[i], 60.0, 6, 10, threads=128)
        col, al = payloads(res[i].avif_file)
        assert col == ref["color_obu"], "colour payload of image %d" % i
        assert ref["alpha_obu"] is not None and al == ref["alpha_obu"], "alpha payload of image %d" % i
    assert enc.encode_rgba(imgs[9]).avif_file == res[9].avif_file
    im = Image.open(io.BytesIO(res[15].avif_file))
    assert im.size == (w, h) and im.mode == "RGBA"


def test_config5_8k_strip_speed2_parity():
    """config #5: an 8192x512 strip of the 8192x8192 layout (1024-wide tiles: 8 of the 64), q90 (quantizer 66), speed 2,
    10-bit: partition range 4..32 searched exhaustively (no early-termination gates), 5/7-type transform search."""
    import cpu_encode as CE
    w, h = 8192, 512
    img = H.synth_image(w, h, 5)
    planes, _, _ = CE.convert(np.ascontiguousarray(img[..., :3]), 10)
    f = CE.configure(w, h, 10, 3, 66, 2, threads=64)
    # the strip's own cap is 4 tiles; the 8192^2 frame has 8 tile columns of 1024: reproduce that column layout
    f8 = H.frame_setup(w, h, 10, 3, 66, 8)
    for k in ("reduced_tx_set", "rdo_tx_decision", "partition_min", "partition_max", "fine_directional_intra", "enable_intra_edge_filter",
              "enable_cdef", "enable_restoration"):
        setattr(f8, k, getattr(f, k))
    for p in range(3):
        f8.lr_type[p] = f.lr_type[p]
    for k in range(3):
        f8.split_gate_bits[k] = f.split_gate_bits[k]
    f = f8
    assert f.tile_cols == 8 and f.tile_rows == 1 and f.partition_max == 32 and f.partition_min == 4
    assert f.split_gate_bits[0] == 0 and f.split_gate_bits[1] == 0
    src = H.pad_planes(planes, f)
    ref = H.oracle_encode(f, H.rate_tables(f.base_q_idx), src)
    ref_fin, ref_deb = H.oracle_filters(f, ref, src)
    out = device_rdo(f, src)
    check_analysis(ref, out, f)
    check_records(ref, out, 3)
    shown = check_filters(f, ref, ref_fin, ref_deb, out, 3, w, h)
    obu = H.pack(f, out)
    dec, bpc, _ = dav1d.decode(obu)
    for p in range(3):
        assert np.array_equal(dec[p], shown[p][:h, :w])


class Av1Config(ctypes.Structure):
    _fields_ = [("width", ctypes.c_size_t), ("height", ctypes.c_size_t), ("bit_depth", ctypes.c_int), ("quantizer", ctypes.c_int),
                ("speed", ctypes.c_int), ("threads", ctypes.c_size_t), ("pixel_range", ctypes.c_int), ("monochrome", ctypes.c_int),
                ("has_color_description", ctypes.c_int), ("matrix_coefficients", ctypes.c_int)]


@pytest.mark.parametrize("depth,speed,quality,threads", [(8, 4, 80.0, 7), (10, 6, 60.0, 3)])
def test_encode_planes_seam_equals_encode_rgb_payload(depth, speed, quality, threads):
    """encode_to_av1 seam (av1encoder.rs:749-771): config + planes in -> OBU bytes out; fed with the planes encode_rgb derives
    (BT.601 full range, ColorDescription{SRGB, BT709, BT601}, :407-411) it must return encode_rgb's primary item."""
    import cpu_encode as CE
    m = C()
    L = H.hostlib()
    w, h = 333, 203
    img = np.ascontiguousarray(H.synth_image(w, h, 21)[..., :3])
    r = m.Encoder.new().with_quality(quality).with_speed(speed).with_bit_depth(depth).with_num_threads(threads).encode_rgb(img)
    col, al = payloads(r.avif_file)
    assert al is None
    planes, _, _ = CE.convert(img, depth)
    # strided planes: the seam takes per-plane strides in samples
    stride = w + 13
    padded = [np.zeros((h, stride), np.uint16) for _ in range(3)]
    for p in range(3):
        padded[p][:, :w] = planes[p]
    cfg = Av1Config(w, h, depth, m.quality_to_quantizer(quality), speed, threads, 1, 0, 1, 6)
    obu = ctypes.c_void_p(); n = ctypes.c_size_t()
    L.b200av1_encode_planes.argtypes = [ctypes.c_void_p] * 5
    rc = L.b200av1_encode_planes(ctypes.byref(cfg), H.ptr3(padded), (ctypes.c_size_t * 3)(stride, stride, stride),
                                 ctypes.byref(obu), ctypes.byref(n))
    assert rc == 0, L.b200avif_last_error_message()
    data = ctypes.string_at(obu.value, n.value)
    libc = ctypes.CDLL(None); libc.free.argtypes = [ctypes.c_void_p]; libc.free(obu)
    assert data == col
    dec, bpc, _ = dav1d.decode(data)
    assert bpc == depth and dec[0].shape == (h, w)
    # monochrome through the same seam (the alpha frame, Cs400, no colour description :443-445)
    cfg = Av1Config(w, h, depth, 121, speed, threads, 1, 1, 0, 2)
    rc = L.b200av1_encode_planes(ctypes.byref(cfg), H.ptr3([padded[0]] * 3), (ctypes.c_size_t * 3)(stride, stride, stride),
                                 ctypes.byref(obu), ctypes.byref(n))
    assert rc == 0, L.b200avif_last_error_message()
    mono = ctypes.string_at(obu.value, n.value); libc.free(obu)
    ref_obu, _, _ = CE.encode_frame([planes[0]], w, h, depth, 121, speed, threads, False)
    assert mono == ref_obu


def test_exif_end_to_end():
    """with_exif (av1encoder.rs:206-219): the item is in the file, libavif returns it, payloads are unchanged."""
    m = C()
    img = H.synth_image(160, 96, 8, alpha=True)
    exif = b"II*\x00\x08\x00\x00\x00" + b"\x01\x00" + b"\x31\x01\x02\x00\x04\x00\x00\x00b20\x00" + b"\x00\x00\x00\x00"
    enc = m.Encoder.new().with_quality(70).with_speed(6)
    plain = enc.encode_rgba(img)
    tagged = enc.with_exif(exif).encode_rgba(img)
    assert payloads(plain.avif_file) == payloads(tagged.avif_file)[:2]
    assert len(tagged.avif_file) > len(plain.avif_file) and exif in tagged.avif_file
    im = Image.open(io.BytesIO(tagged.avif_file))
    assert im.size == (160, 96) and im.mode == "RGBA"
    assert bytes(im.info["exif"]).endswith(exif) and im.getexif().get(0x0131) == "b20"
    assert Image.open(io.BytesIO(plain.avif_file)).info.get("exif") is None


def test_wide_frame_8_mod_16_every_region_decided():
    """regression: coded size 8 mod 16 with more root regions than n_mi / 16 + 64 (4000x3000-class frames)"""
    import cpu_encode as CE
    m = C()
    for (w, h) in ((2056, 8), (4104, 24)):
        img = H.synth_image(w, h, w)
        r = m.Encoder.new().with_quality(75).with_speed(4).with_bit_depth(8).with_num_threads(4).encode_rgba(img)
        ref = CE.encode_rgba(img, 75.0, 4, 8, threads=4)
        assert payloads(r.avif_file)[0] == ref["color_obu"]
        assert Image.open(io.BytesIO(r.avif_file)).size == (w, h)
