"""GPU parity tests (run on the B200 box): every device stage called through the C ABI is compared
BIT-EXACTLY with the CPU oracle on the same seeded inputs, and the emitted bitstream is decoded by
dav1d and compared sample-exactly with the device reconstruction."""
import ctypes
import io
import numpy as np
import pytest
import dav1d
import harness as H

pytestmark = pytest.mark.gpu


def _dev():
    L = H.hostlib()
    assert L.b200avif_device_count() > 0, "CUDA device required: the product has no CPU fallback"
    return L


def device_rdo(f, src):
    L = _dev()
    out = H.Records(f)
    out.deblocked = [np.zeros_like(r) for r in out.rec]
    out.final = [np.zeros_like(r) for r in out.rec]
    src3 = list(src) + [src[0]] * (3 - len(src))
    out.lr_out = [np.zeros_like(r) for r in out.rec]
    out.lf_level = np.zeros(4, np.int32)
    fd = H.device_frame(f)          # the oracle may already have replaced f's deblocking levels by the searched ones
    rc = L.b200av1_rdo_run(ctypes.byref(fd), H.ptr3(src3), out.mi.ctypes.data, H.ptr3(out.eob), H.ptr3(out.coef), H.ptr3(out.rec),
                           H.ptr3(out.deblocked), H.ptr3(out.final), out.cdef_idx.ctypes.data, H.ptrn(out.dec), H.ptrn(out.satd),
                           H.ptr3(out.lr), H.ptr3(out.lr_out), out.lf_level.ctypes.data)
    assert rc == 0, L.b200avif_last_error_message()
    return out


def check_analysis(ref, out, f):
    """stage A: SATD tables and per-block decisions (cost, modes, transform type) of every analysed level"""
    for lv in range(H.LEVELS):
        n = 4 << lv
        if n < min(f.partition_min, 8) or n > f.partition_max:
            continue
        assert np.array_equal(ref.satd[lv], out.satd[lv]), "SATD table level %d" % lv
        for k in ("cost", "y_mode", "tx_type", "uv_mode", "all_zero", "angle_y", "angle_uv", "cfl_js", "cfl_alpha"):
            assert np.array_equal(ref.dec[lv][k], out.dec[lv][k]), "decision %s level %d" % (k, lv)


def check_records(ref, out, planes):
    for p in range(planes):
        assert np.array_equal(ref.rec[p], out.rec[p]), "reconstruction plane %d" % p
        assert np.array_equal(ref.eob[p], out.eob[p]), "eob plane %d" % p
        assert np.array_equal(ref.coef[p], out.coef[p]), "coefficients plane %d" % p
    assert np.array_equal(ref.mi, out.mi), "mode info"


def check_filters(f, ref, ref_fin, ref_deb, out, planes, w, h):
    """in-loop filter stage: deblocked planes, CDEF index + output, restoration units + restored planes.
    Returns the planes a decoder shows."""
    uses_lr = any(f.lr_type[p] for p in range(planes))
    ref_cdef = ref.cdef_out if uses_lr else ref_fin
    for p in range(planes):
        assert np.array_equal(out.deblocked[p], ref_deb[p]), "deblocked plane %d" % p
        assert np.array_equal(out.final[p], ref_cdef[p]), "CDEF output plane %d" % p
    assert np.array_equal(out.cdef_idx, ref.cdef_idx)
    assert [int(v) for v in out.lf_level] == [int(v) for v in f.lf_level], "deblocking levels (search)"
    if not uses_lr:
        return out.final
    for p in range(planes):
        assert np.array_equal(out.lr[p], ref.lr[p]), "restoration units plane %d" % p
        assert np.array_equal(out.lr_out[p][:h, :w], ref_fin[p][:h, :w]), "restored plane %d" % p
    return out.lr_out


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("color_model", [0, 1])
def test_convert_kernel_bit_exact(depth, color_model):
    L = _dev()
    O = H.oracle()
    rng = np.random.default_rng(depth + color_model)
    w, h, stride = 333, 131, 340
    buf = rng.integers(0, 256, (h, stride, 4), dtype=np.uint8)
    buf[:7, :9, :3] = [[255, 0, 0]]; buf[7:9, :9, :3] = [[0, 0, 255]]      # the saturating corner cases
    cw, ch = 8 * ((w + 7) // 8), 8 * ((h + 7) // 8)
    dev = [np.zeros((ch, cw), np.uint16) for _ in range(4)]
    flag = ctypes.c_int(0)
    rc = L.b200av1_convert_run(buf.ctypes.data, 4, w, h, stride, depth, color_model,
                               (ctypes.c_void_p * 4)(*[d.ctypes.data for d in dev]), ctypes.byref(flag))
    assert rc == 0, L.b200avif_last_error_message()
    ref = [np.zeros((h, w), np.uint16) for _ in range(4)]
    O.orc_convert_image(buf.ctypes.data, 4, w, h, stride, color_model, depth, *[r.ctypes.data for r in ref])
    for p in range(4):
        assert np.array_equal(dev[p][:h, :w], ref[p]), "plane %d" % p
        assert np.array_equal(dev[p][:h, w:], np.repeat(ref[p][:, -1:], cw - w, 1))      # edge replication
        assert np.array_equal(dev[p][h:, :w], np.repeat(ref[p][-1:, :], ch - h, 0))
    assert flag.value == 1
    buf[..., 3] = 255
    L.b200av1_convert_run(buf.ctypes.data, 4, w, h, stride, depth, color_model,
                          (ctypes.c_void_p * 4)(*[d.ctypes.data for d in dev]), ctypes.byref(flag))
    assert flag.value == 0


@pytest.mark.parametrize("seed,w,h,stride", [(0, 97, 61, 97), (1, 333, 131, 340), (2, 1, 1, 1), (3, 5, 300, 8)])
def test_alpha_clean_kernels_match_oracle(seed, w, h, stride):
    """dirtyalpha.rs on the device (alpha_clean.cu) vs the oracle restatement: UnassociatedClean and Premultiplied"""
    L = H.hostlib(); O = H.oracle()
    rng = np.random.default_rng(seed)
    img = np.zeros((h, stride, 4), np.uint8)
    img[:, :w] = H.synth_image(w, h, seed, alpha=True)
    # sprinkle every alpha class next to each other, including a = 1 (the u8 truncation case of premultiplied_minmax)
    k = rng.integers(0, 4, (h, w))
    img[:, :w, 3] = np.where(k == 0, 0, np.where(k == 1, 255, np.where(k == 2, 1, img[:, :w, 3])))
    ref = np.zeros((h, w, 4), np.uint8); out = np.zeros((h, w, 4), np.uint8)
    changed = O.orc_blurred_dirty_alpha(img.ctypes.data, w, h, stride, ref.ctypes.data)
    if not changed:
        ref = img[:, :w].copy()
    assert L.b200av1_alpha_clean_run(img.ctypes.data, w, h, stride, 1, out.ctypes.data) == 0, L.b200avif_last_error_message()
    assert np.array_equal(ref, out)
    O.orc_premultiplied_branch(img.ctypes.data, w, h, stride, ref.ctypes.data)
    assert L.b200av1_alpha_clean_run(img.ctypes.data, w, h, stride, 2, out.ctypes.data) == 0
    assert np.array_equal(ref, out)
    # an opaque image is passed through untouched (blurred_dirty_alpha() -> None)
    img[..., 3] = 255
    assert O.orc_blurred_dirty_alpha(img.ctypes.data, w, h, stride, ref.ctypes.data) == 0
    assert L.b200av1_alpha_clean_run(img.ctypes.data, w, h, stride, 1, out.ctypes.data) == 0
    assert np.array_equal(img[:, :w], out)


@pytest.mark.parametrize("depth", [8, 10])
def test_convert_exhaustive_2_24_colours(depth):
    L = _dev(); O = H.oracle()
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    dev = [np.zeros((4096, 4096), np.uint16) for _ in range(3)]
    rc = L.b200av1_convert_run(img.ctypes.data, 3, 4096, 4096, 4096, depth, 0,
                               (ctypes.c_void_p * 4)(*([d.ctypes.data for d in dev] + [None])), None)
    assert rc == 0, L.b200avif_last_error_message()
    ref = [np.zeros((4096, 4096), np.uint16) for _ in range(3)]
    O.orc_convert_image(img.ctypes.data, 3, 4096, 4096, 4096, 0, depth, ref[0].ctypes.data, ref[1].ctypes.data, ref[2].ctypes.data, None)
    for p in range(3):
        assert np.array_equal(dev[p], ref[p])


CASES = [
    # w, h, depth, planes, quantizer, tiles, overrides
    (64, 64, 8, 3, 121, 1, {}),
    (72, 40, 8, 3, 121, 1, {}),
    (129, 101, 10, 3, 181, 1, dict(partition_min=16, rdo_tx_decision=0)),
    (200, 136, 10, 1, 121, 4, {}),
    (640, 360, 8, 3, 147, 12, dict(partition_min=8, rdo_tx_decision=0, reduced_tx_set=0)),
    (512, 512, 10, 3, 66, 16, {}),
    (256, 200, 8, 3, 199, 2, {}),
    (129, 101, 10, 3, 181, 1, dict(partition_min=16, partition_max=16, rdo_tx_decision=0, enable_cdef=0)),
    (640, 360, 8, 3, 121, 4, dict(split_gates=(44, 96))),      # early-termination gates of the fast presets
    (384, 256, 10, 3, 147, 2, dict(split_gates=(44, 96), partition_min=8)),
    (200, 136, 8, 3, 121, 1, dict(lr=1)),                      # loop restoration: merged last unit row
    (161, 100, 10, 3, 147, 2, dict(lr=1)),                     # odd sizes, two tiles
    (95, 95, 8, 1, 100, 1, dict(lr=1)),                        # one unit larger than 64 x 64, monochrome
    (640, 360, 10, 3, 66, 4, dict(lr=1, split_gates=(44, 96))),
    (256, 256, 8, 3, 180, 1, dict(partition_max=32)),           # 32x32 level (speed <= 2)
    (200, 136, 10, 3, 150, 2, dict(partition_max=32, lr=1)),
    (384, 320, 8, 3, 220, 4, dict(partition_max=32, partition_min=8)),
    (200, 136, 8, 3, 121, 2, dict(enable_intra_edge_filter=1)),         # intra edge filter / upsampling
    (256, 192, 10, 3, 100, 1, dict(enable_intra_edge_filter=1, partition_max=32, lr=1)),
    (136, 104, 8, 3, 150, 1, dict(enable_intra_edge_filter=1, y_mode_mask=0x1fe)),       # directional luma modes only
    (320, 200, 8, 3, 100, 2, dict(reduced_tx_set=0)),
    (640, 360, 8, 3, 121, 4, dict(psy=1)),                                               # Tune::Psychovisual luma error
    (320, 256, 10, 3, 66, 2, dict(psy=1, partition_max=32, reduced_tx_set=0, lr=1, lf_search=1)),
    (640, 360, 8, 3, 121, 4, dict(lf_search=1, lr=1)),                                   # deblocking level search (every 4th superblock row and column is sampled)
    (384, 256, 10, 3, 180, 2, dict(lf_search=1, partition_min=8)),
    (200, 136, 8, 1, 100, 1, dict(lf_search=1, lr=1, lr_full=1)),                        # all 16 self-guided sets                                    # 7-type transform set (V_DCT / H_DCT on 4x4, 8x8)
    (200, 136, 10, 3, 66, 1, dict(reduced_tx_set=0, partition_max=32, tx_type_mask=0x61)),
    (320, 200, 8, 3, 90, 2, dict(luma_rdo_modes=4, psy=1, reduced_tx_set=0)),           # complex_prediction_modes: four fully evaluated luma modes
    (200, 136, 10, 3, 66, 1, dict(luma_rdo_modes=4, partition_max=32, psy=1, lr=1)),
]


@pytest.mark.parametrize("w,h,depth,planes,q,tiles,kw", CASES)
def test_rdo_kernel_matches_oracle_and_dav1d(w, h, depth, planes, q, tiles, kw):
    f = H.frame_setup(w, h, depth, planes, q, tiles)
    for k, v in kw.items():
        if k == "split_gates":
            f.split_gate_bits[0], f.split_gate_bits[1] = v
        elif k == "lr":
            f.enable_restoration = 1
            for p in range(planes):
                f.lr_type[p] = 1
        else:
            setattr(f, k, v)
    rt = H.rate_tables(f.base_q_idx)
    img = H.synth_image(w, h, w + h)
    src = H.pad_planes([img[..., c].astype(np.uint16) * (4 if depth == 10 else 1) for c in range(planes)], f)
    ref = H.oracle_encode(f, rt, src)
    ref_fin, ref_deb = H.oracle_filters(f, ref, src)
    out = device_rdo(f, src)
    check_analysis(ref, out, f)
    check_records(ref, out, planes)
    shown = check_filters(f, ref, ref_fin, ref_deb, out, planes, w, h)
    obu = H.pack(f, out)
    dec, bpc, _ = dav1d.decode(obu)     # independent decoder, all in-loop filters on
    for p in range(planes):
        assert np.array_equal(dec[p], shown[p][:h, :w])


@pytest.mark.parametrize("rows", [1, 2, 3, 4])
@pytest.mark.parametrize("depth,pmax", [(8, 16), (10, 32)])
def test_reconstruction_chain_row_slots(rows, depth, pmax, monkeypatch):
    """every wavefront depth of the reconstruction chain (superblock-row slots per tile, normally picked from the shared-memory
    budget) gives the oracle's records: one tile of six superblock rows, two tile columns"""
    monkeypatch.setenv("B200AVIF_RECON_ROWS", str(rows))
    w, h, planes = 200, 328, 3
    f = H.frame_setup(w, h, depth, planes, 90, 2)
    f.partition_max = pmax
    f.partition_min = 4
    rt = H.rate_tables(f.base_q_idx)
    img = H.synth_image(w, h, 77 + depth)
    src = H.pad_planes([img[..., c].astype(np.uint16) * (4 if depth == 10 else 1) for c in range(planes)], f)
    ref = H.oracle_encode(f, rt, src)
    out = device_rdo(f, src)
    check_records(ref, out, planes)


@pytest.mark.parametrize("speed,quantizer", [(6, 121), (4, 121), (8, 66)])
def test_alpha_frame_of_the_fast_presets(speed, quantizer):
    """the monochrome alpha frame exactly as encode_rgba() configures it (real alpha plane: large flat areas, heavy pruning)"""
    import sys, os
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import cpu_encode as CE
    w, h = 360, 240
    img = H.synth_image(w, h, 7, alpha=True)
    _, alpha, anya = CE.convert(img, 10)
    assert anya
    f = CE.configure(w, h, 10, 1, quantizer, speed)
    f.full_range = 1
    src = H.pad_planes([alpha], f)
    ref = H.oracle_encode(f, H.rate_tables(f.base_q_idx), src)
    ref_fin, ref_deb = H.oracle_filters(f, ref, src)
    out = device_rdo(f, src)
    check_analysis(ref, out, f)
    check_records(ref, out, 1)
    check_filters(f, ref, ref_fin, ref_deb, out, 1, w, h)
    assert H.pack(f, out) == H.pack(f, ref)


def test_rdo_full_hd_frame_parity():
    """BASELINE.json config #2 geometry (1920x1080, q80 -> quantizer 121, speed 4, 8-bit, 31 tiles)."""
    import sys, os
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import cpu_encode as CE
    w, h = 1920, 1080
    img = H.synth_image(w, h, 0)
    planes, _, _ = CE.convert(img, 8)
    f = CE.configure(w, h, 8, 3, 121, 4, threads=64)
    assert f.tile_cols * f.tile_rows >= 31
    src = H.pad_planes(planes, f)
    ref = H.oracle_encode(f, H.rate_tables(f.base_q_idx), src)
    ref_fin, ref_deb = H.oracle_filters(f, ref, src)
    out = device_rdo(f, src)
    check_analysis(ref, out, f)
    check_records(ref, out, 3)
    check_filters(f, ref, ref_fin, ref_deb, out, 3, w, h)
