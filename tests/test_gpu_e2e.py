"""End to end through the public API (the call a ravif user makes), on the B200 box.
Restates the reference's own tests (ravif/src/lib.rs:43-147, tests/stdio.rs) against this encoder and
checks the produced files byte-for-byte against the CPU pipeline (oracle hot path + host packer)."""
import io, os, sys
import numpy as np
import pytest
from PIL import Image
import harness as H

sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
pytestmark = pytest.mark.gpu


def C():
    import cavif_rs_b200 as m
    assert m.device_count() > 0
    return m


def payloads(avif):
    """(primary item bytes, alpha item bytes or None) by walking iloc; enough for this muxer's layout."""
    i = avif.find(b"iloc")
    n = int.from_bytes(avif[i + 10:i + 12], "big")
    items = {}
    p = i + 12
    for _ in range(n):
        iid = int.from_bytes(avif[p:p + 2], "big")
        off = int.from_bytes(avif[p + 6:p + 10], "big"); ln = int.from_bytes(avif[p + 10:p + 14], "big")
        items[iid] = avif[off:off + ln]
        p += 14
    return items.get(1), items.get(2)


def test_encode8_with_alpha():            # ravif/src/lib.rs:43-69
    m = C()
    y, x = np.mgrid[0:200, 0:256]
    img = np.stack([x & 255, y & 255, np.full_like(x, 255), (x + y) & 255], -1).astype(np.uint8)
    enc = (m.Encoder.new().with_quality(22.0).with_bit_depth(m.BitDepth.Eight).with_speed(1).with_alpha_quality(22.0)
           .with_alpha_color_mode(m.AlphaColorMode.UnassociatedDirty).with_num_threads(2))
    r = enc.encode_rgba(img)
    assert 50 < r.color_byte_size < 1000 and 50 < r.alpha_byte_size < 1000
    im = Image.open(io.BytesIO(r.avif_file))
    assert im.size == (256, 200) and im.mode == "RGBA"
    col, al = payloads(r.avif_file)
    assert 100 < len(col) < 1000 and al is not None


def test_encode8_opaque_and_rgb_equals_rgba():   # ravif/src/lib.rs:71-119
    m = C()
    y, x = np.mgrid[0:101, 0:129]
    img = np.stack([np.full_like(x, 255), (100 + x) & 255, y & 255, np.full_like(x, 255)], -1).astype(np.uint8)
    enc = (m.Encoder.new().with_quality(33.0).with_speed(10).with_alpha_quality(33.0).with_bit_depth(m.BitDepth.Auto)
           .with_alpha_color_mode(m.AlphaColorMode.UnassociatedDirty).with_num_threads(1))
    r1 = enc.encode_rgba(img)
    assert r1.alpha_byte_size == 0 and 150 < r1.color_byte_size < 500
    im = Image.open(io.BytesIO(r1.avif_file)); assert im.size == (129, 101) and im.mode == "RGB"
    r2 = enc.with_bit_depth(m.BitDepth.Ten).encode_rgb(np.ascontiguousarray(img[..., :3]))
    assert r2.alpha_byte_size == 0
    assert payloads(r1.avif_file)[0] == payloads(r2.avif_file)[0]       # both are the same pixels


def test_encode8_cleans_alpha():          # ravif/src/lib.rs:121-147
    m = C()
    y, x = np.mgrid[0:200, 0:256]
    a = np.clip((((x + y) & 0xff) & 0x7f).astype(int) - 100, 0, 255)
    img = np.stack([(((x // 5 + y) & 0xF) << 4) & 255, (7 * x + y // 2) & 255, (x * y) & 3, a], -1).astype(np.uint8)
    enc = (m.Encoder.new().with_quality(66.0).with_speed(6).with_alpha_quality(88.0)
           .with_alpha_color_mode(m.AlphaColorMode.UnassociatedDirty).with_num_threads(1))
    dirty = enc.encode_rgba(img)
    clean = enc.with_alpha_color_mode(m.AlphaColorMode.UnassociatedClean).encode_rgba(img)
    assert clean.alpha_byte_size == dirty.alpha_byte_size
    assert 200 < clean.alpha_byte_size < 1000 and 2000 < clean.color_byte_size < 6000
    assert clean.color_byte_size < dirty.color_byte_size / 2


def test_testimage_like_cli_config_1():   # BASELINE config #1: 128x85 opaque, q80 s4 depth 8
    m = C()
    img = H.synth_image(128, 85, 1)
    r = m.Encoder.new().with_quality(80).with_speed(4).with_bit_depth(m.BitDepth.Eight).with_alpha_quality(90).encode_rgba(img)
    assert r.avif_file[4:12] == b"ftypavif"          # tests/stdio.rs:23
    dec = np.asarray(Image.open(io.BytesIO(r.avif_file)).convert("RGB")).astype(float)
    mse = np.mean((dec - img[..., :3]) ** 2)
    assert 10 * np.log10(255 ** 2 / mse) > 30


@pytest.mark.parametrize("depth,alpha,speed,quality", [(8, False, 4, 80.0), (10, True, 6, 60.0), (10, False, 4, 90.0),
                                                       (10, False, 2, 90.0),      # speed 2: partition range 4..32, exhaustive
                                                       (8, True, 1, 70.0), (8, False, 10, 40.0)])
def test_api_payload_identical_to_cpu_pipeline(depth, alpha, speed, quality):
    """Same input through (device path via C ABI) and (oracle hot path + host packer): identical AV1 payload bytes."""
    import cpu_encode as CE
    m = C()
    img = H.synth_image(360, 240, 7, alpha=alpha)
    enc = m.Encoder.new().with_quality(quality).with_speed(speed).with_bit_depth(depth)
    r = enc.encode_rgba(img)
    ref = CE.encode_rgba(img, quality, speed, depth)
    col, al = payloads(r.avif_file)
    assert col == ref["color_obu"]
    if alpha:
        assert al == ref["alpha_obu"]
    else:
        assert al is None


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (17, 9), (65, 63), (8, 520), (2049, 16), (130, 97)])
def test_odd_sizes_match_cpu_pipeline(w, h):
    """tiny / odd / extreme aspect-ratio frames: padding, single partial superblock, many forced splits"""
    import cpu_encode as CE
    m = C()
    img = H.synth_image(w, h, w * 7 + h, alpha=(w % 2 == 1))
    r = m.Encoder.new().with_quality(72).with_speed(5).encode_rgba(img)
    ref = CE.encode_rgba(img, 72.0, 5, 10)
    col, al = payloads(r.avif_file)
    assert col == ref["color_obu"]
    assert al == ref["alpha_obu"]
    im = Image.open(io.BytesIO(r.avif_file))
    assert im.size == (w, h)


@pytest.mark.parametrize("w,h,speed,quality", [(17, 9, 1, 70.0), (40, 24, 1, 70.0), (200, 136, 2, 70.0), (24, 24, 2, 95.0)])
def test_mixed_partition_ranges_in_one_launch(w, h, speed, quality):
    """speed <= 2: the colour frame (quantizer > 121: blocks up to 16x16) and the alpha frame (up to 32x32) of one image go through
    the same launches with different top levels; frames smaller than a 32x32 block have no block at the top level at all"""
    import cpu_encode as CE
    m = C()
    img = H.synth_image(w, h, w + 3 * h, alpha=True)
    r = m.Encoder.new().with_quality(quality).with_speed(speed).with_bit_depth(8).encode_rgba(img)
    ref = CE.encode_rgba(img, quality, speed, 8)
    col, al = payloads(r.avif_file)
    assert col == ref["color_obu"] and al == ref["alpha_obu"]


def test_premultiplied_mode_matches_cpu_pipeline():
    import cpu_encode as CE
    m = C()
    img = H.synth_image(200, 136, 4, alpha=True)
    r = m.Encoder.new().with_quality(70).with_speed(6).with_alpha_color_mode(m.AlphaColorMode.Premultiplied).encode_rgba(img)
    ref = CE.encode_rgba(img, 70.0, 6, 10, alpha_mode=2)
    col, al = payloads(r.avif_file)
    assert col == ref["color_obu"] and al == ref["alpha_obu"]


def test_concurrent_calls_share_one_encoder():
    """N OS threads call encode_rgba on the same Encoder at once (cavif: rayon par_iter over files, src/main.rs:223)"""
    import threading
    m = C()
    enc = m.Encoder.new().with_quality(65).with_speed(6)
    imgs = [H.synth_image(160 + 16 * i, 96 + 8 * i, 20 + i, alpha=(i % 3 == 0)) for i in range(8)]
    serial = [enc.encode_rgba(im).avif_file for im in imgs]
    out = [None] * len(imgs)
    def work(i):
        out[i] = enc.encode_rgba(imgs[i]).avif_file
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(imgs))]
    for t in th: t.start()
    for t in th: t.join()
    assert out == serial


def test_batch_matches_single_and_isolates_errors():
    m = C()
    imgs = [H.synth_image(200 + 8 * i, 120, i, alpha=(i % 2 == 1)) for i in range(5)]
    enc = m.Encoder.new().with_quality(75).with_speed(5)
    batch = enc.encode_batch(imgs)
    for im, b in zip(imgs, batch):
        assert enc.encode_rgba(im).avif_file == b.avif_file


def test_very_large_batch_in_one_launch(monkeypatch):
    """more frame jobs in ONE launch set than the machine has resident CTAs (640 tiny images with alpha = 1280 jobs, one chunk):
    grids that are sized per job, survivor lists, tile tables and the arenas all scale with the job count"""
    monkeypatch.setenv("B200AVIF_CHUNKS", "1")
    m = C()
    imgs = [H.synth_image(24 + 8 * (i % 3), 16 + 8 * (i % 2), i, alpha=True) for i in range(640)]
    enc = m.Encoder.new().with_quality(70).with_speed(4)
    batch = enc.encode_batch(imgs)
    assert len(batch) == 640
    monkeypatch.delenv("B200AVIF_CHUNKS")
    for i in (0, 1, 2, 317, 638, 639):
        assert enc.encode_rgba(imgs[i]).avif_file == batch[i].avif_file
        im = Image.open(io.BytesIO(batch[i].avif_file))
        assert im.size == (imgs[i].shape[1], imgs[i].shape[0]) and im.mode == "RGBA"


def test_raw_planes_and_errors():
    m = C()
    w, h = 64, 48
    rng = np.random.default_rng(0)
    planes = rng.integers(0, 1024, (w * h, 3)).astype(np.uint16)
    r = m.Encoder.new().with_speed(6).encode_raw_planes_10_bit(w, h, planes, None, m.PixelRange.Full, m.MatrixCoefficients.BT601)
    assert Image.open(io.BytesIO(r.avif_file)).size == (w, h)
    with pytest.raises(m.TooFewPixels):
        m.Encoder.new().encode_raw_planes_10_bit(w, h, planes[:-1], None, m.PixelRange.Full, m.MatrixCoefficients.BT601)
    with pytest.raises(m.Unsupported):
        m.Encoder.new().encode_raw_planes_8_bit(w, h, (planes >> 2).astype(np.uint8), None, m.PixelRange.Full, 4)


def test_strided_input_and_device_resident_input():
    import torch
    m = C()
    big = H.synth_image(320, 200, 9)
    view = big[10:170, 16:272]               # strided sub-image (imgref stride semantics)
    a = m.Encoder.new().with_speed(6).encode_rgba(view)
    b = m.Encoder.new().with_speed(6).encode_rgba(np.ascontiguousarray(view))
    assert a.avif_file == b.avif_file
    t = torch.from_numpy(np.ascontiguousarray(view)).cuda()
    c = m.Encoder.new().with_speed(6).with_alpha_color_mode(m.AlphaColorMode.UnassociatedDirty).encode_rgba(t)
    assert c.avif_file == a.avif_file
